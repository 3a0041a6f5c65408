"""Thin Python driver over the libserenc C ABI: owns the handle, allocates tensors/workspaces with torch
(device memory + streams only) and calls the encode entry points. No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
import threading
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .configs import ARCH_TEXT, ARCH_WHISPER, ARCH_W2V, EncoderConfig, w2v_num_frames

REDUCE_NONE = 0
REDUCE_MEAN = 1
REDUCE_WEIGHTED = 2   # sum_i w_i * hidden_states[sel_i] (lora_wavlm/model.py:164-181 of the reference)


def _wav_dtype(wav: torch.Tensor) -> int:
    if wav.dtype == torch.float32:
        return _lib.WAV_F32
    if wav.dtype == torch.int16:
        return _lib.WAV_I16
    raise TypeError(f"waveforms must be float32 or int16 PCM, got {wav.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def layer_mask_of(indices: Iterable[int], num_layers: int) -> Tuple[int, List[int]]:
    """hidden_states indices (negative allowed, as in `hidden_states[-4:]`) -> (bit mask, sorted indices)."""
    n = num_layers + 1
    idx = sorted({(i + n) if i < 0 else i for i in indices})
    for i in idx:
        if not 0 <= i < n:
            raise IndexError(f"hidden state index {i} out of range for {n} hidden states")
    mask = 0
    for i in idx:
        mask |= 1 << i
    return mask, idx


class UploadRing:
    """Copy/compute overlap for a stream of packed batches (one host thread, two streams per GPU).

    `upload(host)` copies a pinned 1-D fp32 tensor into one of `slots` device buffers on a dedicated copy stream and
    makes the CURRENT stream wait for that copy only. A slot is reused `slots` uploads later, after the compute that
    read it (marked by `release`). Because kernel launches are asynchronous, the host reaches the next `upload` while
    the GPU is still encoding the previous batch, so the transfer of batch i+1 runs under the encode of batch i."""

    def __init__(self, device: torch.device, slots: int = 2):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device)
        self._bufs: List[Optional[torch.Tensor]] = [None] * slots
        self._free: List[Optional[torch.cuda.Event]] = [None] * slots
        self._i = 0
        # pinned host staging, one buffer per slot, grown on demand and then reused: packing a batch costs a memcpy, not a
        # cudaHostAlloc per call (round 1 pinned a fresh tensor for every batch)
        self._hbufs: List[Optional[torch.Tensor]] = [None] * slots
        self._h2d_done: List[Optional[torch.cuda.Event]] = [None] * slots

    def upload_arrays(self, arrays, dtype: torch.dtype) -> Tuple[torch.Tensor, int]:
        """Pack numpy arrays back to back into this slot's pinned staging buffer (converted to `dtype`; int16 PCM that has
        to become float32 is scaled by 1 / 32768 as librosa does) and upload it."""
        import numpy as np
        s = self._i % len(self._bufs)
        n = int(sum(len(a) for a in arrays))
        es = torch.empty((), dtype=dtype).element_size()
        hb = self._hbufs[s]
        if self._h2d_done[s] is not None:
            self._h2d_done[s].synchronize()      # the previous upload out of this host buffer (two batches ago) has been read
        if hb is None or hb.numel() < n * es:
            hb = self._hbufs[s] = torch.empty(max(int(n * es * 1.25), 1 << 20), dtype=torch.uint8).pin_memory()
        host = hb[: n * es].view(dtype)
        dst = host.numpy()
        off = 0
        for a in arrays:
            m = len(a)
            if dtype == torch.float32 and getattr(a, "dtype", None) == np.int16:
                np.multiply(a, np.float32(1.0 / 32768.0), out=dst[off:off + m], dtype=np.float32)
            else:
                dst[off:off + m] = a
            off += m
        out = self.upload(host)
        return out

    def upload(self, host: torch.Tensor) -> Tuple[torch.Tensor, int]:
        assert host.dtype in (torch.float32, torch.int16, torch.int32) and host.dim() == 1 and not host.is_cuda
        if not host.is_pinned():
            host = host.pin_memory()
        s = self._i % len(self._bufs)
        self._i += 1
        n = host.numel()
        cur = torch.cuda.current_stream(self.device)
        buf = self._bufs[s]
        if buf is not None and buf.dtype != host.dtype:    # the slot is a byte pool: re-typed views, no new allocation
            nb = buf.numel() * buf.element_size()
            buf = buf.view(torch.uint8)[: nb - nb % host.element_size()].view(host.dtype)
            self._bufs[s] = buf
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=host.dtype, device=self.device)   # allocated on the compute stream ...
            buf.record_stream(self.copy_stream)                                       # ... and written on the copy stream
            self._bufs[s] = buf
            self.copy_stream.wait_stream(cur)    # the allocator may hand back memory the compute stream is still using
        if self._free[s] is not None:
            self.copy_stream.wait_event(self._free[s])
        with torch.cuda.stream(self.copy_stream):
            buf[:n].copy_(host, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._h2d_done[s] = done
        cur.wait_event(done)
        return buf[:n], s

    def release(self, slot: int) -> None:
        """Everything enqueued on the current stream so far may read the slot; later uploads into it wait for that."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev


class DownloadRing:
    """Frame-level results back to the host without stalling the GPU: the packed [rows, d] result of a
    batch is copied on a copy stream into one of `slots` pinned buffers; the caller gets the pinned view and an event
    and hands both to a writer thread (which waits for the event, clones each utterance's rows and saves them) while
    the main thread launches the next batch. A slot is only reused after `release(slot)`."""

    def __init__(self, device: torch.device, slots: int = 3):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device)
        self._bufs: List[Optional[torch.Tensor]] = [None] * slots
        self._busy = [threading.Event() for _ in range(slots)]
        for e in self._busy:
            e.set()          # set = free
        self._i = 0

    def download(self, packed: torch.Tensor) -> Tuple[torch.Tensor, torch.cuda.Event, int]:
        assert packed.is_cuda and packed.is_contiguous()
        s = self._i % len(self._bufs)
        self._i += 1
        self._busy[s].wait()                       # back-pressure: the writers are `slots` batches behind
        self._busy[s].clear()
        n = packed.numel()
        buf = self._bufs[s]
        if buf is None or buf.numel() < n:
            buf = self._bufs[s] = torch.empty(n, dtype=packed.dtype).pin_memory()
        host = buf[:n].view(packed.shape)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))    # the encode that produces `packed`
        packed.record_stream(self.copy_stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            host.copy_(packed, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        return host, done, s

    def release(self, slot: int) -> None:
        self._busy[slot].set()


class Engine:
    """One encoder replica on one GPU."""

    def __init__(self, cfg: EncoderConfig, tensors: Dict[str, np.ndarray], device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("interspeech_ser_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError("interspeech_ser_b200 runs on CUDA devices only")
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device(f"cuda:{self.device_index}")
        self._lib = _lib.load_library()
        c = _lib.SerencConfig()
        c.arch = cfg.arch
        c.hidden, c.layers, c.heads, c.ffn = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size
        c.conv_dim = cfg.conv_dim[0]
        c.conv_bias = int(cfg.conv_bias)
        c.wavlm_rel_bias = int(cfg.family == "wavlm")
        c.num_buckets, c.max_distance = cfg.num_buckets, cfg.max_bucket_distance
        c.pos_conv_kernel, c.pos_conv_groups = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        c.n_mels, c.max_source_positions = cfg.num_mel_bins, cfg.max_source_positions
        c.layer_norm_eps = cfg.layer_norm_eps
        c.conv_group_norm = int(cfg.feat_extract_norm == "group")
        c.post_layer_norm = int(not cfg.do_stable_layer_norm)
        c.no_feat_proj_ln = int(not cfg.feat_proj_layer_norm)
        c.vocab_size, c.max_positions = cfg.vocab_size, cfg.max_position_embeddings
        c.type_vocab_size, c.pad_token_id = cfg.type_vocab_size, cfg.pad_token_id
        if cfg.arch == ARCH_W2V:
            if cfg.feat_extract_norm not in ("layer", "group"):
                raise ValueError(f"feat_extract_norm={cfg.feat_extract_norm!r} has to be 'layer' or 'group'")
            if tuple(cfg.conv_kernel) != (10, 3, 3, 3, 3, 2, 2) or tuple(cfg.conv_stride) != (5, 2, 2, 2, 2, 2, 2):
                raise NotImplementedError("unsupported feature-encoder geometry")
        h = C.c_void_p()
        torch.cuda.init()
        _lib.check(self._lib.serenc_create(C.byref(c), self.device_index, C.byref(h)))
        self._h = h
        try:
            for name, arr in tensors.items():
                arr = np.ascontiguousarray(arr, dtype=np.float32)
                shape = _lib.i64_array(arr.shape if arr.ndim else (1,))
                _lib.check(self._lib.serenc_load_tensor(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), shape,
                                                        max(arr.ndim, 1)))
            _lib.check(self._lib.serenc_finalize(self._h))
        except Exception:
            self.close()
            raise
        self._tls = threading.local()
        self._graphs: "OrderedDict[tuple, tuple]" = OrderedDict()   # CUDA-graph cache of small encode calls (encode_w2v_graphed)
        self._graph_lock = threading.Lock()

    # ------------------------------------------------------------------ lifecycle
    def synchronize(self) -> None:
        """Wait for the current stream and surface an asynchronous kernel fault as SerencError (the handle is then
        poisoned: every later call fails with the same message, include/serenc.h)."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_sync(self._h, _stream_ptr(self.device)))

    @property
    def poisoned(self) -> bool:
        return bool(self._h) and bool(self._lib.serenc_is_poisoned(self._h))

    def close(self) -> None:
        if getattr(self, "_graphs", None):
            self._graphs.clear()
        if getattr(self, "_h", None):
            self._lib.serenc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, nbytes: int) -> torch.Tensor:
        ws = getattr(self._tls, "ws", None)
        if ws is None or ws.numel() < nbytes:
            self._tls.ws = None
            ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=self.device)
            self._tls.ws = ws
        return ws

    # ------------------------------------------------------------------ accounting
    PROF_CLASSES = ("gemm_linear", "gemm_conv", "gemm_posconv", "attention", "layernorm", "conv0", "pool", "logmel", "misc",
                    "gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2")
    GEMM_CLASSES = ("gemm_linear", "gemm_conv", "gemm_posconv", "gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2")

    def launch_count(self) -> int:
        return int(self._lib.serenc_launch_count(self._h))

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self._lib.serenc_set_profiling(self._h, int(enable)))

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        n = len(self.PROF_CLASSES)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(self._lib.serenc_get_profile(self._h, n, ms, fl, by, cnt))
        return {name: {"ms": ms[i], "flops": fl[i], "bytes": by[i], "launches": int(cnt[i])} for i, name in enumerate(self.PROF_CLASSES)}

    # ------------------------------------------------------------------ wav2vec2 family
    def w2v_workspace_bytes(self, lens: Sequence[int]) -> int:
        out = C.c_size_t()
        _lib.check(self._lib.serenc_w2v_workspace_bytes(self._h, _lib.i32_array(lens), len(lens), C.byref(out)))
        return int(out.value)

    def normalize(self, wav: torch.Tensor, starts: Sequence[int], lens: Sequence[int], out_len: int) -> torch.Tensor:
        """Feature-extractor normalisation on the GPU -> [B, out_len] fp32 (zero right padding)."""
        B = len(lens)
        out = torch.empty((B, out_len), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_wav_normalize(self._h, wav.data_ptr(), _lib.i64_array(starts), _lib.i32_array(lens), B,
                                                      out.data_ptr(), out_len, out_len, _stream_ptr(self.device)))
        return out

    def encode_w2v(self, wav: torch.Tensor, starts: Sequence[int], lens: Sequence[int], *, normalize: bool,
                   layers: Iterable[int], reduce: int = REDUCE_NONE, want_frames: bool = True,
                   want_pooled: bool = False, layer_weights: Optional[Sequence[float]] = None,
                   want_extract_features: bool = False):
        """Returns (frames | None, pooled | None, frame_offsets[B+1], selected layer indices[, extract_features]).
        frames: [n_sel, sum_T, d] (REDUCE_NONE) or [sum_T, d]; pooled: [n_sel, B, d] or [B, d].
        wav: float32 samples or int16 PCM (scaled by 1/32768 inside the kernels, as librosa does on the host).
        layer_weights: one weight per selected layer (ascending index) for REDUCE_WEIGHTED.
        want_extract_features: additionally return HF's `extract_features` [sum_T, conv_dim] fp32 as a fifth value."""
        assert wav.is_cuda and wav.is_contiguous()
        wav_dtype = _wav_dtype(wav)
        B = len(lens)
        mask, idx = layer_mask_of(layers, self.cfg.num_hidden_layers)
        if reduce == REDUCE_WEIGHTED and (layer_weights is None or len(layer_weights) != len(idx)):
            raise ValueError(f"REDUCE_WEIGHTED needs one weight per selected layer ({len(idx)})")
        T = [w2v_num_frames(n, self.cfg) for n in lens]
        for b, t in enumerate(T):
            if t < 1:
                raise ValueError(f"utterance {b}: {lens[b]} samples is shorter than the 400-sample receptive field")
        sumT, d = sum(T), self.cfg.hidden_size
        lead = () if reduce != REDUCE_NONE else (len(idx),)
        frames = torch.empty(lead + (sumT, d), dtype=torch.float32, device=self.device) if want_frames else None
        pooled = torch.empty(lead + (B, d), dtype=torch.float32, device=self.device) if want_pooled else None
        feats = torch.empty((sumT, self.cfg.conv_dim[-1]), dtype=torch.float32, device=self.device) if want_extract_features else None
        ws = self._workspace(self.w2v_workspace_bytes(lens))
        offs = (C.c_int64 * (B + 1))()
        call = _lib.SerencW2VCall()
        call.wav_dev, call.wav_dtype, call.batch = wav.data_ptr(), wav_dtype, B
        call.sample_start, call.sample_len = _lib.i64_array(starts), _lib.i32_array(lens)
        call.normalize, call.reduce, call.layer_mask = int(normalize), reduce, mask
        if reduce == REDUCE_WEIGHTED:
            call.layer_weights = _lib.f32_array(layer_weights)
        call.frames_out_dev, call.pooled_out_dev, call.extract_features_out_dev = _ptr(frames), _ptr(pooled), _ptr(feats)
        call.frame_offsets_out = offs
        call.workspace_dev, call.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            call.stream = _stream_ptr(self.device)
            _lib.check(self._lib.serenc_encode_w2v_ex(self._h, C.byref(call)))
        if want_extract_features:
            return frames, pooled, list(offs), idx, feats
        return frames, pooled, list(offs), idx

    # Small batches are launch-bound (~250 launches of a few microseconds: 8 x 4 s takes 3.3 ms eager, 2.9 ms replayed,
    # profiles/r01_notes.md), so pooled-only calls below this many frames are captured into a CUDA graph once per length
    # signature and replayed; larger batches are not launch-bound and stay eager.
    GRAPH_MAX_FRAMES = 4096
    GRAPH_CACHE_SIZE = 16

    def encode_w2v_graphed(self, wav: torch.Tensor, starts: Sequence[int], lens: Sequence[int], *, normalize: bool,
                           layers: Iterable[int], reduce: int, layer_weights: Optional[Sequence[float]] = None):
        """Pooled-only encode_w2v through a CUDA graph keyed on (lengths, layers, reduce, weights, dtype). The waveform is
        copied into the graph's static input buffer and the pooled result is cloned out of its static output, so the
        call is interchangeable with the eager one (bit-identical results: the same kernels in the same order)."""
        # the stream is part of the key: every stream replays its own graph with its own static buffers
        key = (tuple(int(n) for n in lens), tuple(int(v) for v in starts), tuple(sorted(layers)), int(reduce),
               None if layer_weights is None else tuple(float(v) for v in layer_weights), wav.dtype, bool(normalize),
               _stream_ptr(self.device))
        with self._graph_lock:
            ent = self._graphs.get(key)
            if ent is not None:
                self._graphs.move_to_end(key)
        if ent is None:
            static_in = torch.empty_like(wav)
            static_in.copy_(wav)
            # capture on a side stream with a workspace that belongs to the graph
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            saved_ws = getattr(self._tls, "ws", None)
            self._tls.ws = None
            try:
                with torch.cuda.stream(side):
                    self.encode_w2v(static_in, starts, lens, normalize=normalize, layers=layers, reduce=reduce, want_frames=False,
                                    want_pooled=True, layer_weights=layer_weights)   # warm-up: allocates the workspace
                    side.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        _, pooled, offs, idx = self.encode_w2v(static_in, starts, lens, normalize=normalize, layers=layers, reduce=reduce,
                                                               want_frames=False, want_pooled=True, layer_weights=layer_weights)
                graph_ws = self._tls.ws
            finally:
                self._tls.ws = saved_ws
            torch.cuda.current_stream(self.device).wait_stream(side)
            ent = (graph, static_in, pooled, offs, idx, graph_ws)
            with self._graph_lock:
                self._graphs[key] = ent
                while len(self._graphs) > self.GRAPH_CACHE_SIZE:
                    self._graphs.popitem(last=False)
        graph, static_in, pooled, offs, idx, _ = ent
        with self._graph_lock:   # threads that share a stream: copy-in / replay / copy-out enqueued as one unit
            static_in.copy_(wav, non_blocking=True)
            graph.replay()
            out = pooled.clone()
        return None, out, list(offs), idx

    def unpack(self, packed: torch.Tensor, offsets: Sequence[int], t_max: int) -> torch.Tensor:
        """packed [sum_T, cols] -> HF-shaped [B, t_max, cols] (pad frames zero)."""
        B = len(offsets) - 1
        cols = int(packed.shape[-1])
        out = torch.empty((B, t_max, cols), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_unpack_rows(self._h, packed.data_ptr(), _lib.i64_array(offsets), B, t_max, cols,
                                                    out.data_ptr(), _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ whisper
    def logmel(self, wav: torch.Tensor, starts: Sequence[int], lens: Sequence[int]) -> torch.Tensor:
        B = len(lens)
        mel = torch.empty((B, self.cfg.num_mel_bins, 3000), dtype=torch.float32, device=self.device)
        scratch = torch.empty(64 + 40 * B + 64, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_logmel_ex(self._h, wav.data_ptr(), _wav_dtype(wav), _lib.i64_array(starts), _lib.i32_array(lens), B,
                                                  mel.data_ptr(), scratch.data_ptr(), _stream_ptr(self.device)))
        return mel

    def whisper_workspace_bytes(self, batch: int) -> int:
        out = C.c_size_t()
        _lib.check(self._lib.serenc_whisper_workspace_bytes(self._h, batch, C.byref(out)))
        return int(out.value)

    def encode_whisper(self, mel: torch.Tensor, *, layers: Iterable[int], reduce: int = REDUCE_NONE,
                       n_keep: Optional[Sequence[int]] = None, want_frames: bool = True, want_pooled: bool = False,
                       layer_weights: Optional[Sequence[float]] = None):
        assert mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous()
        if mel.dim() != 3 or mel.shape[1] != self.cfg.num_mel_bins or mel.shape[2] != 3000:
            # HF WhisperEncoder.forward raises ValueError here (modeling_whisper.py:613-617)
            raise ValueError(f"Whisper expects the mel input features to be of length 3000, but found {tuple(mel.shape)}. "
                             "Make sure to pad the input mel features to 3000.")
        B = mel.shape[0]
        mask, idx = layer_mask_of(layers, self.cfg.num_hidden_layers)
        d = self.cfg.hidden_size
        if reduce == REDUCE_WEIGHTED and (layer_weights is None or len(layer_weights) != len(idx)):
            raise ValueError(f"REDUCE_WEIGHTED needs one weight per selected layer ({len(idx)})")
        lead = () if reduce != REDUCE_NONE else (len(idx),)
        frames = torch.empty(lead + (B * 1500, d), dtype=torch.float32, device=self.device) if want_frames else None
        pooled = torch.empty(lead + (B, d), dtype=torch.float32, device=self.device) if want_pooled else None
        ws = self._workspace(self.whisper_workspace_bytes(B))
        keep = _lib.i32_array(n_keep) if n_keep is not None else None
        lw = _lib.f32_array(layer_weights) if reduce == REDUCE_WEIGHTED else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_encode_whisper_ex(self._h, mel.data_ptr(), B, mask, reduce, lw, keep, _ptr(frames),
                                                          _ptr(pooled), ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        return frames, pooled, idx

    # ------------------------------------------------------------------ text encoder
    def text_workspace_bytes(self, batch: int, seq_len: int) -> int:
        out = C.c_size_t()
        _lib.check(self._lib.serenc_text_workspace_bytes(self._h, batch, seq_len, C.byref(out)))
        return int(out.value)

    def encode_text(self, input_ids: torch.Tensor, valid_len: Sequence[int], *, layers: Iterable[int], reduce: int = REDUCE_NONE,
                    want_frames: bool = True, want_pooled: bool = False, layer_weights: Optional[Sequence[float]] = None):
        """input_ids: [B, T] int32 on the device, right-padded; valid_len[b] = non-pad tokens of row b.
        Returns (frames [n_sel, B*T, d] | [B*T, d] | None, pooled | None, selected layer indices)."""
        assert input_ids.is_cuda and input_ids.dtype == torch.int32 and input_ids.is_contiguous() and input_ids.dim() == 2
        B, T = input_ids.shape
        mask, idx = layer_mask_of(layers, self.cfg.num_hidden_layers)
        if reduce == REDUCE_WEIGHTED and (layer_weights is None or len(layer_weights) != len(idx)):
            raise ValueError(f"REDUCE_WEIGHTED needs one weight per selected layer ({len(idx)})")
        d = self.cfg.hidden_size
        lead = () if reduce != REDUCE_NONE else (len(idx),)
        frames = torch.empty(lead + (B * T, d), dtype=torch.float32, device=self.device) if want_frames else None
        pooled = torch.empty(lead + (B, d), dtype=torch.float32, device=self.device) if want_pooled else None
        ws = self._workspace(self.text_workspace_bytes(B, T))
        lw = _lib.f32_array(layer_weights) if reduce == REDUCE_WEIGHTED else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.serenc_encode_text(self._h, input_ids.data_ptr(), _lib.i32_array(valid_len), B, T, mask, reduce, lw,
                                                    _ptr(frames), _ptr(pooled), ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        return frames, pooled, idx
