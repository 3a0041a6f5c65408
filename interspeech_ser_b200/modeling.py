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
from .configs import ARCH_WHISPER, EncoderConfig, get_config, w2v_num_frames
from .engine import REDUCE_MEAN, REDUCE_NONE, Engine, UploadRing
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

    def _select(self, layer: int, average: bool) -> Tuple[List[int], int]:
        if average:  # torch.stack(hidden_states[-4:]) — Python slicing: fewer than 4 states means "all of them"
            n = self.cfg.num_hidden_layers + 1
            return list(range(max(0, n - 4), n)), REDUCE_MEAN
        return [layer], REDUCE_NONE


class SpeechEncoderModel(_Base):
    """WavLMModel / Wav2Vec2Model / HubertModel replacement."""

    def forward(self, input_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                output_hidden_states: Optional[bool] = None, output_attentions: Optional[bool] = None,
                return_dict: Optional[bool] = None, mask_time_indices=None, **kw) -> ModelOutput:
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash-style kernel")
        if input_values.dim() == 1:
            input_values = input_values[None]
        x = input_values.to(self.device, torch.float32).contiguous()
        B, Lmax = x.shape
        if attention_mask is not None:
            lens = [int(v) for v in attention_mask.to(torch.float32).ne(0).sum(-1).tolist()]
        else:
            lens = [Lmax] * B
        starts = [b * Lmax for b in range(B)]
        L = self.cfg.num_hidden_layers
        layers = range(L + 1) if output_hidden_states else [L]
        frames, _, offs, idx = self.engine.encode_w2v(x, starts, lens, normalize=False, layers=layers,
                                                      reduce=REDUCE_NONE, want_frames=True, want_pooled=False)
        t_max = w2v_num_frames(Lmax, self.cfg)
        if B == 1:
            hs = tuple(frames[i].view(1, t_max, -1) for i in range(len(idx)))
        else:
            hs = tuple(self.engine.unpack(frames[i], offs, t_max) for i in range(len(idx)))
        return ModelOutput(last_hidden_state=hs[-1], extract_features=None, hidden_states=hs if output_hidden_states else None)

    __call__ = forward

    @torch.no_grad()
    def extract(self, waveforms: Sequence[np.ndarray], layer: int = -1, average: bool = False, want_frames: bool = True,
                want_pooled: bool = True) -> Extracted:
        """Batched embedding extraction: raw 16 kHz waveforms -> hidden_states[layer] (or mean of the last four,
        preprocess_speech.py:56-63) per utterance + masked-mean pooled vectors. One H2D copy, one encode call."""
        lens = [int(len(w)) for w in waveforms]
        flat = torch.from_numpy(np.concatenate([np.asarray(w, dtype=np.float32) for w in waveforms]))
        return self.extract_pinned(flat.pin_memory(), lens, layer=layer, average=average, want_frames=want_frames, want_pooled=want_pooled)

    @torch.no_grad()
    def extract_device(self, wav: torch.Tensor, lens: Sequence[int], layer: int = -1, average: bool = False,
                       want_frames: bool = False, want_pooled: bool = True) -> Extracted:
        """Same as extract() for a packed waveform tensor already resident on the device."""
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        layers, reduce = self._select(layer, average)
        frames, pooled, offs, _ = self.engine.encode_w2v(wav, starts, lens, normalize=self.cfg.do_normalize, layers=layers,
                                                         reduce=reduce, want_frames=want_frames, want_pooled=want_pooled)
        per_utt, f2, ranges = None, None, None
        if frames is not None:
            f2 = frames if reduce == REDUCE_MEAN else frames[0]
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
                want_pooled: bool = True, literal_crop: bool = True) -> Extracted:
        """log-mel -> encoder -> hidden_states[layer] | mean of last four -> keep the first
        min(ceil(len/320), cap) frames (preprocess_whisper.py:49-50,75-76; cap = hidden size when literal_crop,
        reproducing the script's `feats.shape[1]`, else 1500)."""
        waveforms = [np.asarray(w, dtype=np.float32)[:480000] for w in waveforms]
        lens = [int(len(w)) for w in waveforms]
        flat = torch.from_numpy(np.concatenate(waveforms)).pin_memory()
        return self.extract_pinned(flat, lens, layer=layer, average=average, want_frames=want_frames, want_pooled=want_pooled,
                                   literal_crop=literal_crop)

    @torch.no_grad()
    def extract_device(self, wav: torch.Tensor, lens: Sequence[int], layer: int = -1, average: bool = False,
                       want_frames: bool = False, want_pooled: bool = True, literal_crop: bool = True) -> Extracted:
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        mel = self.engine.logmel(wav, starts, lens)
        cap = self.cfg.hidden_size if literal_crop else 1500
        keep = [max(1, min(-(-n // 320), cap, 1500)) for n in lens]
        layers, reduce = self._select(layer, average)
        frames, pooled, _ = self.engine.encode_whisper(mel, layers=layers, reduce=reduce, n_keep=keep, want_frames=want_frames,
                                                       want_pooled=want_pooled)
        per_utt, f2, ranges = None, None, None
        if frames is not None:
            f2 = frames if reduce == REDUCE_MEAN else frames[0]
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
        model = (WhisperModel if cfg.arch == ARCH_WHISPER else SpeechEncoderModel)(cfg, tensors, device)
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
