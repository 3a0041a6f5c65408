"""Processor objects with the call surface the reference scripts use.

    ssl_processor = AutoFeatureExtractor.from_pretrained(SSL_TYPE)                  preprocess_speech.py:111
    inputs = ssl_processor(y, sampling_rate=sr, return_tensors="pt", padding=True)  preprocess_speech.py:48
    ssl_processor = AutoProcessor.from_pretrained(SSL_TYPE)                         preprocess_whisper.py:119
    inputs = ssl_processor(y, sampling_rate=sr, return_tensors="pt")["input_features"]   preprocess_whisper.py:48-53

The arithmetic (zero-mean/unit-variance normalisation; log-mel spectrogram) runs in libserenc on the GPU; the
returned tensors therefore live on the processor's CUDA device (`.to(device)` in the scripts is then a no-op).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .configs import EncoderConfig

AudioLike = Union[np.ndarray, torch.Tensor, Sequence[float]]


class BatchFeature(dict):
    """dict with attribute access and `.to(device)`, like transformers.BatchFeature."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        return BatchFeature({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})


def _as_list(raw) -> List[np.ndarray]:
    if torch.is_tensor(raw):
        raw = raw.detach().cpu().numpy()
    if isinstance(raw, np.ndarray):
        if raw.ndim == 1:
            return [np.ascontiguousarray(raw, dtype=np.float32)]
        if raw.ndim == 2:
            return [np.ascontiguousarray(r, dtype=np.float32) for r in raw]
        raise ValueError("Only mono-channel audio is supported")
    if isinstance(raw, (list, tuple)) and len(raw) and isinstance(raw[0], (np.ndarray, list, tuple)) or \
            (isinstance(raw, (list, tuple)) and len(raw) and torch.is_tensor(raw[0])):
        return [np.ascontiguousarray(r.detach().cpu().numpy() if torch.is_tensor(r) else r, dtype=np.float32) for r in raw]
    return [np.ascontiguousarray(raw, dtype=np.float32)]


def _check_sr(given: Optional[int], expected: int, cls: str) -> None:
    # HF raises ValueError on a mismatching rate (feature_extraction_wav2vec2.py:168-174)
    if given is not None and given != expected:
        raise ValueError(f"The model corresponding to this feature extractor: {cls} was trained using a sampling rate of "
                         f"{expected}. Please make sure that the provided `raw_speech` input was sampled with {expected} "
                         f"and not {given}.")


class Wav2Vec2FeatureExtractor:
    """Wav2Vec2FeatureExtractor(do_normalize=True, return_attention_mask=True) surface
    (HF models/wav2vec2/feature_extraction_wav2vec2.py:99-236)."""

    model_input_names = ["input_values", "attention_mask"]

    def __init__(self, cfg: EncoderConfig, engine=None):
        self.cfg = cfg
        self.sampling_rate = cfg.sampling_rate
        self.do_normalize = cfg.do_normalize
        self.return_attention_mask = cfg.return_attention_mask
        self.padding_value = 0.0
        self._engine = engine

    def bind(self, engine) -> "Wav2Vec2FeatureExtractor":
        self._engine = engine
        return self

    def __call__(self, raw_speech: AudioLike, sampling_rate: Optional[int] = None, return_tensors: Optional[str] = "pt",
                 padding: Union[bool, str] = False, return_attention_mask: Optional[bool] = None, **kwargs) -> BatchFeature:
        _check_sr(sampling_rate, self.sampling_rate, type(self).__name__)
        if return_tensors not in ("pt", None):
            raise ValueError("only return_tensors='pt' is supported")
        if self._engine is None:
            raise RuntimeError("feature extractor is not bound to a GPU engine (use AutoFeatureExtractor.from_pretrained "
                               "after AutoModel.from_pretrained, or .bind(model)); there is no CPU path")
        waves = _as_list(raw_speech)
        lens = [int(len(w)) for w in waves]
        if len(waves) > 1 and not padding and len(set(lens)) > 1:
            raise ValueError("Unable to create tensor: utterances have different lengths, pass padding=True")
        lmax = max(lens)
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        eng = self._engine
        flat = torch.from_numpy(np.concatenate(waves)).pin_memory().to(eng.device, non_blocking=True)
        if self.do_normalize:
            values = eng.normalize(flat, starts, lens, lmax)
        else:
            values = torch.zeros((len(waves), lmax), dtype=torch.float32, device=eng.device)
            for b, (s, n) in enumerate(zip(starts, lens)):
                values[b, :n] = flat[s:s + n]
        out = {"input_values": values}
        want_mask = self.return_attention_mask if return_attention_mask is None else return_attention_mask
        if want_mask:
            mask = torch.zeros((len(waves), lmax), dtype=torch.int32)
            for b, n in enumerate(lens):
                mask[b, :n] = 1
            out["attention_mask"] = mask.to(eng.device)
        return BatchFeature(out)


class WhisperFeatureExtractor:
    """WhisperFeatureExtractor surface (HF models/whisper/feature_extraction_whisper.py:189-342): pads/truncates to
    30 s and returns {"input_features": [B, n_mels, 3000]} computed by the log-mel kernel."""

    model_input_names = ["input_features"]

    def __init__(self, cfg: EncoderConfig, engine=None):
        self.cfg = cfg
        self.sampling_rate = cfg.sampling_rate
        self.feature_size = cfg.num_mel_bins
        self.n_fft, self.hop_length, self.chunk_length = 400, 160, 30
        self.n_samples = self.chunk_length * self.sampling_rate
        self.nb_max_frames = self.n_samples // self.hop_length
        self._engine = engine

    def bind(self, engine) -> "WhisperFeatureExtractor":
        self._engine = engine
        return self

    def __call__(self, raw_speech: AudioLike, sampling_rate: Optional[int] = None, return_tensors: Optional[str] = "pt",
                 **kwargs) -> BatchFeature:
        _check_sr(sampling_rate, self.sampling_rate, type(self).__name__)
        if self._engine is None:
            raise RuntimeError("processor is not bound to a GPU engine; there is no CPU path")
        waves = [w[: self.n_samples] for w in _as_list(raw_speech)]
        lens = [int(len(w)) for w in waves]
        starts, off = [], 0
        for n in lens:
            starts.append(off)
            off += n
        eng = self._engine
        flat = torch.from_numpy(np.concatenate(waves)).pin_memory().to(eng.device, non_blocking=True)
        return BatchFeature({"input_features": eng.logmel(flat, starts, lens)})


class WhisperProcessor(WhisperFeatureExtractor):
    """AutoProcessor.from_pretrained("openai/whisper-*") returns a processor whose audio path is the feature
    extractor; the tokenizer half is not on this path."""

    @property
    def feature_extractor(self):
        return self
