"""Model objects with the HuggingFace call surface the reference scripts use, backed by libserenc.

    ssl_model = AutoModel.from_pretrained(SSL_TYPE); ssl_model.eval(); ssl_model.to(device)   preprocess_speech.py:112-114
    outputs = model(**inputs, output_hidden_states=True)                                      preprocess_speech.py:50-54
    outputs.hidden_states  /  outputs['hidden_states'][N]                                     preprocess_speech.py:56,67
    model.encoder(input_features, output_hidden_states=True)                                  preprocess_whisper.py:57,71
    ssl_model(x, attention_mask=mask).last_hidden_state ; .config.hidden_size ; .freeze_feature_encoder()
                                                                  benchmark/train_eval_files/train_cat_ser.py:118-123,173-175

Beyond the drop-in surface, `extract()` is the batched fast path the CLIs and the benchmark use: raw waveforms in,
selected / averaged hidden states and masked-mean pooled embeddings out, normalisation fused into the first kernel.
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import weights as W
from .configs import ARCH_TEXT, ARCH_WHISPER, EncoderConfig, get_config, w2v_num_frames
from .engine import REDUCE_MEAN, REDUCE_NONE, REDUCE_WEIGHTED, Engine, UploadRing
from .feature_extraction import Wav2Vec2FeatureExtractor, WhisperFeatureExtractor, WhisperProcessor


class ModelOutput(OrderedDict):
    """Attribute + key + index access, like transformers.utils.ModelOutput (None fields are skipped)."""

    def __init__(self, **kw):
        super().__init__((k, v) for k, v in kw.items() if v is not None)
        self._all = kw

    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        allv = self.__dict__.get("_all", {})
        if k in allv:
            return allv[k]
        raise AttributeError(k)

    def __getitem__(self, k):
        if isinstance(k, (int, slice)):
            return tuple(self.values())[k]
        return super().__getitem__(k)

    def to_tuple(self):
        return tuple(self.values())


class _HFConfigView:
    """The few `model.config.*` attributes downstream code reads."""

    def __init__(self, cfg: EncoderConfig):
        self._cfg = cfg
        self.hidden_size = cfg.hidden_size
        self.d_model = cfg.hidden_size
        self.num_hidden_layers = cfg.num_hidden_layers
        self.encoder_layers = cfg.num_hidden_layers
        self.num_attention_heads = cfg.num_attention_heads
        self.intermediate_size = cfg.intermediate_size
        self.num_mel_bins = cfg.num_mel_bins
        self.max_source_positions = cfg.max_source_positions
        self.model_type = cfg.family
        self.output_hidden_states = False

    def to_dict(self):
        return self._cfg.to_dict()


@dataclass
class Extracted:
    """Result of the batched fast path."""
    frames: Optional[List[torch.Tensor]]   # per utterance [T_keep, d] fp32 (views into one packed device tensor)
    pooled: Optional[torch.Tensor]         # [B, d] fp32 masked-mean embeddings
    num_frames: List[int]
    packed: Optional[torch.Tensor] = None  # the [rows, d] device tensor `frames` are views of (one D2H moves them all)
    ranges: Optional[List[Tuple[int, int]]] = None   # row range of every utterance inside `packed`


class _Base:
    def __init__(self, cfg: EncoderConfig, tensors: Dict[str, np.ndarray], device=0):
        self.cfg = cfg
        self.config = _HFConfigView(cfg)
        self.engine = Engine(cfg, tensors, device)
        self.device = self.engine.device
        self.training = False
        self._ring_tls = threading.local()   # one upload ring per calling thread (the CLI's workers share the model)

    def _upload(self, host: torch.Tensor):
        ring = getattr(self._ring_tls, "ring", None)
        if ring is None:
            ring = self._ring_tls.ring = UploadRing(self.device)
        wav, slot = ring.upload(host)
        return ring, wav, slot

    @torch.no_grad()
    def _extract_arrays(self, waveforms: Sequence[np.ndarray], lens: Sequence[int], **kw) -> "Extracted":
        """extract() for numpy waveforms: packed straight into the upload ring's reusable pinned staging buffer (int16 when
        every input is int16 PCM, else float32), uploaded on the copy stream, encoded."""
        all_i16 = len(waveforms) > 0 and all(getattr(w, "dtype", None) == np.int16 for w in waveforms)
        with torch.cuda.device(self.device):
            ring = getattr(self._ring_tls, "ring", None)
            if ring is None:
                ring = self._ring_tls.ring = UploadRing(self.device)
            wav, slot = ring.upload_arrays(waveforms, torch.int16 if all_i16 else torch.float32)
            res = self.extract_device(wav, lens, **kw)
            ring.release(slot)
        return res

    @torch.no_grad()
    def extract_pinned(self, host: torch.Tensor, lens: Sequence[int], **kw) -> "Extracted":
        """extract() for utterances already packed back to back in one pinned host tensor: the upload goes through a
        two-slot ring on a copy stream, so in a loop over batches it overlaps the previous batch's encode."""
        with torch.cuda.device(self.device):
            ring, wav, slot = self._upload(host.reshape(-1))
            res = self.extract_device(wav, lens, **kw)
            ring.release(slot)
        return res

    # nn.Module-ish no-ops the scripts call
    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("inference-only encoder")
        return self

    def to(self, device=None, *a, **k):
        if device is not None:
            dev = torch.device(device)
            if dev.type != "cuda":
                raise RuntimeError("this encoder has no CPU path; it lives on the CUDA device it was created on")
            if dev.index is not None and dev.index != self.device.index:
                raise RuntimeError(f"encoder was created on {self.device}; create a second replica for {dev}")
        return self

    def cuda(self, device=None):
        return self.to(f"cuda:{device}" if isinstance(device, int) else (device or self.device))

    def requires_grad_(self, flag: bool = False):
        return self

    def parameters(self):
        return iter(())

    def freeze_feature_encoder(self):
        return None

    def _select(self, layer: int, average: bool, layer_weights: Optional[Sequence[float]] = None,
                layers: Optional[Sequence[int]] = None) -> Tuple[List[int], int, Optional[List[float]]]:
        """-> (hidden-state indices, reduce mode, per-layer weights in ascending layer order)."""
        n = self.cfg.num_hidden_layers + 1
        if layer_weights is not None:
            # softmax-weighted layer sum (lora_wavlm/model.py:164-181): weights for `layers` (default: all but the conv
            # output, the reference's use_conv_output = False); the caller has applied the softmax
            sel = list(layers) if layers is not None else list(range(n - len(layer_weights), n))
            sel = [(i + n) if i < 0 else i for i in sel]
            if len(sel) != len(layer_weights) or len(set(sel)) != len(sel):
                raise ValueError("layer_weights needs one weight per distinct selected layer")
            order = sorted(range(len(sel)), key=lambda j: sel[j])
            return [sel[j] for j in order], REDUCE_WEIGHTED, [float(layer_weights[j]) for j in order]
        if average:  # torch.stack(hidden_states[-4:]) — Python slicing: fewer than 4 states means "all of them"
            return list(range(max(0, n - 4), n)), REDUCE_MEAN, None
        return [layer], REDUCE_NONE, None

    # ---- request coalescing for the reference's unchanged 4-thread call pattern (preprocess_speech.py:120-122) ----
    def enable_request_batching(self, max_batch: int = 64, max_wait_ms: float = 2.0) -> "_Base":
        """Concurrent `model(**inputs)` calls (one utterance each, as the reference's ThreadPoolExecutor issues them) are
        collected by a dispatcher thread and encoded as ONE packed batch; every caller gets exactly the result the
        batch-1 call would have produced (batching invariance is bit-exact, tests/test_gpu_models.py)."""
        if getattr(self, "_queue", None) is None:
            self._queue = _CoalescingQueue(self, max_batch, max_wait_ms)
        return self

    def disable_request_batching(self) -> None:
        q = getattr(self, "_queue", None)
        if q is not None:
            q.close()
            self._queue = None


class _CoalescingQueue:
    """Dispatcher thread behind enable_request_batching(). A request is (run, args): `collect` turns the requests that
    are pending at the same time into one engine call and hands every caller its slice."""

    def __init__(self, model: "_Base", max_batch: int, max_wait_ms: float):
        import queue

        self.model = model
        self.max_batch = max(1, int(max_batch))
        self.max_wait = max(0.0, float(max_wait_ms)) / 1e3
        self.q: "queue.Queue" = queue.Queue()
        self.batches = 0          # statistics: engine calls issued / requests served
        self.requests = 0
        self._closed = False
        self._thread = threading.Thread(target=self._loop, name="serenc-batcher", daemon=True)
        self._thread.start()

    def submit(self, item):
        from concurrent.futures import Future

        if self._closed:
            raise RuntimeError("request batching was disabled")
        fut: Future = Future()
        self.q.put((item, fut))
        return fut

    def close(self):
        self._closed = True
        self.q.put(None)
        self._thread.join(timeout=5.0)

    def _loop(self):
        import queue
        import time

        while True:
            first = self.q.get()
            if first is None:
                return
            group = [first]
            deadline = time.monotonic() + self.max_wait
            while len(group) < self.max_batch:
                try:
                    nxt = self.q.get(timeout=max(0.0, deadline - time.monotonic()))
                except queue.Empty:
                    break
                if nxt is None:
                    self.q.put(None)
                    break
                group.append(nxt)
            # requests that ask for different outputs cannot share a call
            by_kind: Dict[object, list] = {}
            for item, fut in group:
                by_kind.setdefault(item[0], []).append((item, fut))
            for kind, members in by_kind.items():
                try:
                    results = self.model._run_coalesced(kind, [it for it, _ in members])
                    self.batches += 1
                    self.requests += len(members)
                    for (_, fut), res in zip(members, results):
                        fut.set_result(res)
                except BaseException as e:  # noqa: BLE001 - every waiting caller must be released
                    for _, fut in members:
                        if not fut.done():
                            fut.set_exception(e)


class SpeechEncoderModel(_Base):
    """WavLMModel / Wav2Vec2Model / HubertModel replacement."""

    def forward(self, input_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                output_hidden_states: Optional[bool] = None, output_attentions: Optional[bool] = None,
                return_dict: Optional[bool] = None, mask_time_indices=None, **kw) -> ModelOutput:
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash-style kernel")
        if input_values.dim() == 1:
            input_values = input_values[None]
        B, Lmax = input_values.shape
        if attention_mask is not None:
            lens = [int(v) for v in attention_mask.to(torch.float32).ne(0).sum(-1).tolist()]
        else:
            lens = [Lmax] * B
        q = getattr(self, "_queue", None)
        if q is not None and threading.current_thread() is not q._thread:
            # one utterance per call from several threads: coalesced into one packed batch by the dispatcher
            item = (bool(output_hidden_states), input_values, lens)
            return q.submit(item).result()
        return self._forward_batch(input_values, lens, bool(output_hidden_states))

    __call__ = forward

    def _forward_batch(self, input_values: torch.Tensor, lens: Sequence[int], all_states: bool) -> ModelOutput:
        x = input_values.to(self.device, torch.float32).contiguous()
        B, Lmax = x.shape
        starts = [b * Lmax for b in range(B)]
        L = self.cfg.num_hidden_layers
        layers = range(L + 1) if all_states else [L]
        want_feats = self.cfg.feat_proj_layer_norm and self.cfg.family != "hubert"   # HubertModel returns no extract_features
        res = self.engine.encode_w2v(x, starts, lens, normalize=False, layers=layers, reduce=REDUCE_NONE, want_frames=True,
                                     want_pooled=False, want_extract_features=want_feats)
        frames, _, offs, idx = res[:4]
        t_max = w2v_num_frames(Lmax, self.cfg)
        if B == 1 and offs[1] == t_max:     # a single full-length row: the packed matrix IS the padded one
            unpack = lambda t: t.view(1, t_max, -1)  # noqa: E731
        else:
            unpack = lambda t: self.engine.unpack(t, offs, t_max)  # noqa: E731
        hs = tuple(unpack(frames[i]) for i in range(len(idx)))
        feats = unpack(res[4]) if want_feats else None
        return ModelOutput(last_hidden_state=hs[-1], extract_features=feats, hidden_states=hs if all_states else None)

    def _run_coalesced(self, all_states: bool, items) -> List[ModelOutput]:
        """Requests collected by the dispatcher -> one packed encode -> one HF-shaped output per request."""
        rows, lens, owner = [], [], []
        for ri, (_, iv, ls) in enumerate(items):
            iv = iv.to(self.device, torch.float32)
            for b, n in enumerate(ls):
                rows.append(iv[b, :n])
                lens.append(n)
                owner.append(ri)
        flat = torch.cat(rows).contiguous()
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        L = self.cfg.num_hidden_layers
        layers = range(L + 1) if all_states else [L]
        want_feats = self.cfg.feat_proj_layer_norm and self.cfg.family != "hubert"
        res = self.engine.encode_w2v(flat, starts, lens, normalize=False, layers=layers, reduce=REDUCE_NONE, want_frames=True,
                                     want_pooled=False, want_extract_features=want_feats)
        frames, _, offs, idx = res[:4]
        outs: List[ModelOutput] = []
        u = 0
        for ri, (_, iv, ls) in enumerate(items):
            B, Lmax = iv.shape
            t_max = w2v_num_frames(Lmax, self.cfg)
            sub = [o - offs[u] for o in offs[u:u + B + 1]]
            a, e = offs[u], offs[u + B]

            def unpack(t, a=a, e=e, sub=sub, B=B, t_max=t_max):
                if B == 1 and sub[1] == t_max:
                    return t[a:e].view(1, t_max, -1)
                return self.engine.unpack(t[a:e].contiguous(), sub, t_max)
            hs = tuple(unpack(frames[i]) for i in range(len(idx)))
            outs.append(ModelOutput(last_hidden_state=hs[-1], extract_features=unpack(res[4]) if want_feats else None,
                                    hidden_states=hs if all_states else None))
            u += B
        return outs

    @torch.no_grad()
    def extract(self, waveforms: Sequence[np.ndarray], layer: int = -1, average: bool = False, want_frames: bool = True,
                want_pooled: bool = True, **kw) -> Extracted:
        """Batched embedding extraction: raw 16 kHz waveforms -> hidden_states[layer] (or mean of the last four,
        preprocess_speech.py:56-63) per utterance + masked-mean pooled vectors. One H2D copy, one encode call.
        Waveforms may be float32 samples or int16 PCM (all of one kind): int16 halves the upload and is scaled by
        1 / 32768 inside the first kernel, exactly as librosa scales it on the host."""
        lens = [int(len(w)) for w in waveforms]
        return self._extract_arrays(waveforms, lens, layer=layer, average=average, want_frames=want_frames,
                                    want_pooled=want_pooled, **kw)

    @torch.no_grad()
    def extract_device(self, wav: torch.Tensor, lens: Sequence[int], layer: int = -1, average: bool = False,
                       want_frames: bool = False, want_pooled: bool = True, layer_weights: Optional[Sequence[float]] = None,
                       layers: Optional[Sequence[int]] = None, use_graph: Optional[bool] = None) -> Extracted:
        """Same as extract() for a packed waveform tensor already resident on the device.
        layer_weights (+ optional layers): weighted sum of hidden states instead of one layer / the mean of the last four.
        use_graph: replay small pooled-only calls from a CUDA graph cached per length signature (default: automatic)."""
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        sel, reduce, lw = self._select(layer, average, layer_weights, layers)
        graph_ok = (want_pooled and not want_frames and reduce != REDUCE_NONE
                    and sum(w2v_num_frames(n, self.cfg) for n in lens) <= self.engine.GRAPH_MAX_FRAMES
                    and not torch.cuda.is_current_stream_capturing())
        if use_graph is None:
            use_graph = graph_ok and os.environ.get("SERENC_NO_GRAPH") != "1"
        if use_graph and graph_ok:
            frames, pooled, offs, _ = self.engine.encode_w2v_graphed(wav, starts, lens, normalize=self.cfg.do_normalize, layers=sel,
                                                                     reduce=reduce, layer_weights=lw)
        else:
            frames, pooled, offs, _ = self.engine.encode_w2v(wav, starts, lens, normalize=self.cfg.do_normalize, layers=sel,
                                                             reduce=reduce, want_frames=want_frames, want_pooled=want_pooled,
                                                             layer_weights=lw)
        per_utt, f2, ranges = None, None, None
        if frames is not None:
            f2 = frames if reduce != REDUCE_NONE else frames[0]
            ranges = [(offs[b], offs[b + 1]) for b in range(len(lens))]
            per_utt = [f2[a:e] for a, e in ranges]
        if pooled is not None and reduce == REDUCE_NONE:
            pooled = pooled[0]
        return Extracted(per_utt, pooled, [offs[b + 1] - offs[b] for b in range(len(lens))], f2, ranges)


class WhisperEncoder:
    """`model.encoder` of WhisperModel (HF modeling_whisper.py:541-647)."""

    def __init__(self, owner: "WhisperModel"):
        self._o = owner
        self.config = owner.config

    def forward(self, input_features: torch.Tensor, attention_mask=None, output_hidden_states: Optional[bool] = None,
                output_attentions: Optional[bool] = None, return_dict: Optional[bool] = None, **kw) -> ModelOutput:
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash-style kernel")
        o = self._o
        mel = input_features.to(o.device, torch.float32).contiguous()
        L = o.cfg.num_hidden_layers
        layers = range(L + 1) if output_hidden_states else [L]
        frames, _, idx = o.engine.encode_whisper(mel, layers=layers, reduce=REDUCE_NONE, want_frames=True, want_pooled=False)
        B = mel.shape[0]
        hs = tuple(frames[i].view(B, 1500, -1) for i in range(len(idx)))
        return ModelOutput(last_hidden_state=hs[-1], hidden_states=hs if output_hidden_states else None)

    __call__ = forward


class WhisperModel(_Base):
    """WhisperModel replacement: only `.encoder` exists (the reference never touches the decoder on this path)."""

    def __init__(self, cfg: EncoderConfig, tensors, device=0):
        super().__init__(cfg, tensors, device)
        self.encoder = WhisperEncoder(self)

    def get_encoder(self):
        return self.encoder

    @torch.no_grad()
    def extract(self, waveforms: Sequence[np.ndarray], layer: int = -1, average: bool = False, want_frames: bool = True,
                want_pooled: bool = True, literal_crop: bool = True, **kw) -> Extracted:
        """log-mel -> encoder -> hidden_states[layer] | mean of last four -> keep the first
        min(ceil(len/320), cap) frames (preprocess_whisper.py:49-50,75-76; cap = hidden size when literal_crop,
        reproducing the script's `feats.shape[1]`, else 1500). float32 samples or int16 PCM."""
        waveforms = [np.asarray(w)[:480000] for w in waveforms]
        lens = [int(len(w)) for w in waveforms]
        return self._extract_arrays(waveforms, lens, layer=layer, average=average, want_frames=want_frames,
                                    want_pooled=want_pooled, literal_crop=literal_crop, **kw)

    @torch.no_grad()
    def extract_device(self, wav: torch.Tensor, lens: Sequence[int], layer: int = -1, average: bool = False,
                       want_frames: bool = False, want_pooled: bool = True, literal_crop: bool = True,
                       layer_weights: Optional[Sequence[float]] = None, layers: Optional[Sequence[int]] = None,
                       use_graph: Optional[bool] = None) -> Extracted:
        # use_graph is accepted for signature parity with SpeechEncoderModel and ignored: a 30 s window is never launch-bound
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        mel = self.engine.logmel(wav, starts, lens)
        cap = self.cfg.hidden_size if literal_crop else 1500
        keep = [max(1, min(-(-n // 320), cap, 1500)) for n in lens]
        sel, reduce, lw = self._select(layer, average, layer_weights, layers)
        frames, pooled, _ = self.engine.encode_whisper(mel, layers=sel, reduce=reduce, n_keep=keep, want_frames=want_frames,
                                                       want_pooled=want_pooled, layer_weights=lw)
        per_utt, f2, ranges = None, None, None
        if frames is not None:
            f2 = frames if reduce != REDUCE_NONE else frames[0]
            ranges = [(b * 1500, b * 1500 + keep[b]) for b in range(len(lens))]
            per_utt = [f2[a:e] for a, e in ranges]
        if pooled is not None and reduce == REDUCE_NONE:
            pooled = pooled[0]
        return Extracted(per_utt, pooled, keep, f2, ranges)


# --------------------------------------------------------------------------------------------------
# Auto* entry points
# --------------------------------------------------------------------------------------------------
_LAST_MODEL: Dict[str, _Base] = {}


def _resolve_weights(cfg: EncoderConfig, name_or_path: str, random_init: Optional[bool], seed: int):
    cand = []
    if os.path.isdir(name_or_path) or name_or_path.endswith((".npz", ".pt", ".pth", ".bin", ".safetensors")):
        cand.append(name_or_path)   # with config_name=...: e.g. a LoRA-tuned classifier state dict (weights.merge_lora)
    root = os.environ.get("SERENC_WEIGHTS_DIR")
    if root:
        cand.append(os.path.join(root, name_or_path))
        cand.append(os.path.join(root, name_or_path.split("/")[-1]))
    for c in cand:
        if os.path.exists(c):
            return W.load_checkpoint_dir(cfg, c)
    if random_init or (random_init is None and os.environ.get("SERENC_RANDOM_INIT") == "1"):
        return W.random_init(cfg, seed)
    # what the reference reports when from_pretrained cannot find the model (preprocess_speech.py:115-117)
    raise OSError(f"No pretrained weights found for '{name_or_path}' (looked in {cand or 'nowhere'}); there is no network in "
                  "this environment. Point SERENC_WEIGHTS_DIR at a directory of HF checkpoints or pass random_init=True.")


class AutoModel:
    @staticmethod
    def from_pretrained(name_or_path: str, device: Union[int, str, torch.device, None] = None,
                        random_init: Optional[bool] = None, seed: int = 0, config_name: Optional[str] = None, **kw):
        cfg = get_config(config_name or name_or_path)
        tensors = _resolve_weights(cfg, name_or_path, random_init, seed)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else 0
        if cfg.arch == ARCH_TEXT:
            from .text import RobertaModel
            model = RobertaModel(cfg, tensors, device)
        else:
            model = (WhisperModel if cfg.arch == ARCH_WHISPER else SpeechEncoderModel)(cfg, tensors, device)
        if os.environ.get("SERENC_BATCH_REQUESTS") == "1" and cfg.arch not in (ARCH_TEXT, ARCH_WHISPER):
            model.enable_request_batching()
        _LAST_MODEL[cfg.name] = model
        return model


def _bound_engine(cfg: EncoderConfig, model):
    if model is not None:
        return model.engine
    m = _LAST_MODEL.get(cfg.name)
    return m.engine if m is not None else None


class AutoFeatureExtractor:
    @staticmethod
    def from_pretrained(name_or_path: str, model=None, config_name: Optional[str] = None, **kw):
        cfg = get_config(config_name or name_or_path)
        eng = _bound_engine(cfg, model)
        if cfg.arch == ARCH_WHISPER:
            return WhisperFeatureExtractor(cfg, eng)
        return Wav2Vec2FeatureExtractor(cfg, eng)


class AutoProcessor:
    @staticmethod
    def from_pretrained(name_or_path: str, model=None, config_name: Optional[str] = None, **kw):
        cfg = get_config(config_name or name_or_path)
        eng = _bound_engine(cfg, model)
        if cfg.arch == ARCH_WHISPER:
            return WhisperProcessor(cfg, eng)
        return Wav2Vec2FeatureExtractor(cfg, eng)
