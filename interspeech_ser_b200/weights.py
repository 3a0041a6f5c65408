"""Weight handling: HF state-dict -> canonical libserenc tensor names, random init, checkpoint files.

Canonical names (the table INTEGRATION.md documents; consumed by serenc_load_tensor):

  wav2vec2 / HuBERT / WavLM                                  Whisper encoder
  -------------------------                                  ---------------
  conv{0..6}.weight|bias|ln.weight|ln.bias                   conv1.weight|bias, conv2.weight|bias
  featproj.ln.weight|bias, featproj.weight|bias              embed_positions
  posconv.weight (weight-norm folded) | posconv.bias         mel_filters   (201 x n_mels, host table)
  rel_attn_embed                (WavLM)
  layer{i}.ln1.*  layer{i}.{q,k,v,o}.*  layer{i}.ln2.*  layer{i}.fc1.*  layer{i}.fc2.*    (both)
  layer{i}.gru.weight|bias|const (WavLM)
  final_ln.weight|bias                                       (both)

  RoBERTa (text branch): embed.word, embed.position, embed.type, final_ln.* (= embeddings.LayerNorm, the stack's
  leading LayerNorm), layer{i}.* as above with ln1 = attention.output.LayerNorm, ln2 = output.LayerNorm.

All tensors are fp32 numpy arrays on the host; dtype/layout conversion for the tensor cores happens inside
the library at load time.
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Mapping

import numpy as np

from .configs import ARCH_TEXT, ARCH_WHISPER, EncoderConfig


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def fold_weight_norm(g, v) -> np.ndarray:
    """nn.utils.weight_norm(conv, dim=2) as used by *PositionalConvEmbedding (HF modeling_wavlm.py:60-76):
    W[:, :, k] = g[k] * v[:, :, k] / ||v[:, :, k]||_F  (norm over dims 0 and 1 for every tap k)."""
    g = _np(g).astype(np.float64).reshape(1, 1, -1)
    v = _np(v).astype(np.float64)
    norm = np.sqrt((v * v).sum(axis=(0, 1), keepdims=True))
    return np.ascontiguousarray((v * (g / norm)).astype(np.float32))


def slaney_mel_filters(n_mels: int, n_freq: int = 201, sr: int = 16000, fmin: float = 0.0,
                       fmax: float = 8000.0) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular mel filterbank [n_freq, n_mels]
    (restates HF audio_utils.mel_filter_bank :453-544 with hertz_to_mel/mel_to_hertz :263-332 and
    _create_triangular_filter_bank :356-375 for the arguments WhisperFeatureExtractor passes)."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        mels = 3.0 * f / 200.0
        log_region = f >= 1000.0
        safe = np.where(log_region, f, 1000.0)
        return np.where(log_region, 15.0 + np.log(safe / 1000.0) * (27.0 / np.log(6.4)), mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = 200.0 * m / 3.0
        log_region = m >= 15.0
        return np.where(log_region, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)

    mel_pts = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
    filter_freqs = mel_to_hz(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    fb = fb * enorm[None, :]
    return np.ascontiguousarray(fb.astype(np.float32))


def whisper_sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Whisper's fixed position table (HF modeling_whisper.py:55-64): cat[sin, cos] of log-spaced timescales."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.ascontiguousarray(np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32))


_LORA_RE = re.compile(r"^(.*)\.lora_([AB])\.([^.]+)\.weight$")


def merge_lora(sd: Mapping[str, object], lora_alpha: float = 16.0, adapter: str = "default") -> Dict[str, object]:
    """Fold peft LoRA adapters into the dense weights:  W <- W + (lora_alpha / r) * B @ A.

    The reference's `*_pretrained` extraction loads a `WavLMClassifier` state dict whose encoder was wrapped by
    `get_peft_model(LoraConfig(r=8, lora_alpha=16, target_modules=['q_proj', 'v_proj']))`
    (preprocessing/preprocess_speech_pretrained.py:120-130) and runs `ssl_model.wavlm.model` in eval mode, where the
    adapter contributes exactly that low-rank term (dropout is off).  peft's key layout is
    `<wrapper>.base_model.model.<module>.base_layer.weight|bias`, `<module>.lora_A.<adapter>.weight [r, in]`,
    `<module>.lora_B.<adapter>.weight [out, r]`; the classifier head and other non-encoder keys are dropped.
    r is read from the adapter's shape; lora_alpha is not stored in a state dict, so the reference's 16 is the default."""
    plain: Dict[str, object] = {}
    lora: Dict[str, Dict[str, object]] = {}
    for k, v in sd.items():
        if "base_model.model." in k:
            k = k.split("base_model.model.", 1)[1]
        elif k.startswith("classifier."):
            continue
        m = _LORA_RE.match(k)
        if m:
            if m.group(3) == adapter:
                lora.setdefault(m.group(1), {})[m.group(2)] = v
            continue
        plain[k.replace(".base_layer.", ".")] = v
    for mod, ab in lora.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"LoRA adapter for '{mod}' is missing lora_{'B' if 'A' in ab else 'A'}")
        a, b = _np(ab["A"]).astype(np.float64), _np(ab["B"]).astype(np.float64)
        if a.shape[0] != b.shape[1]:
            raise ValueError(f"LoRA rank mismatch for '{mod}': A {a.shape} vs B {b.shape}")
        key = mod + ".weight"
        if key not in plain:
            raise ValueError(f"LoRA adapter for '{mod}' has no base weight")
        w = _np(plain[key]).astype(np.float64)
        if w.shape != (b.shape[0], a.shape[1]):
            raise ValueError(f"LoRA shape mismatch for '{mod}': W {w.shape}, B@A {(b.shape[0], a.shape[1])}")
        plain[key] = (w + (float(lora_alpha) / a.shape[0]) * (b @ a)).astype(np.float32)
    return plain


def from_hf_state_dict(cfg: EncoderConfig, sd: Mapping[str, object], lora_alpha: float = 16.0) -> Dict[str, np.ndarray]:
    """Convert a HuggingFace state dict (WavLMModel / Wav2Vec2Model / HubertModel / WhisperModel or
    WhisperEncoder, with or without a task-head prefix, with or without peft LoRA adapters) into canonical tensors."""
    sd = dict(sd)
    if any(".lora_A." in k for k in sd):
        sd = merge_lora(sd, lora_alpha)
    # strip wrapper prefixes: "wavlm.", "wav2vec2.", "hubert.", "model." (and peft's "whisper." / "roberta." task wrappers)
    for prefix in ("wavlm.", "wav2vec2.", "hubert.", "whisper.", "roberta.", "model."):
        if any(k.startswith(prefix) for k in sd):
            sd = {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}
    out: Dict[str, np.ndarray] = {}
    L = cfg.num_hidden_layers
    if cfg.arch == ARCH_TEXT:
        # RobertaModel (HF modeling_roberta.py): RobertaEmbeddings + RobertaLayer x L; the pooler is not on this path
        e = "embeddings."
        out["embed.word"] = _np(sd[e + "word_embeddings.weight"])
        out["embed.position"] = _np(sd[e + "position_embeddings.weight"])
        out["embed.type"] = _np(sd[e + "token_type_embeddings.weight"])
        out["final_ln.weight"] = _np(sd[e + "LayerNorm.weight"])
        out["final_ln.bias"] = _np(sd[e + "LayerNorm.bias"])
        for i in range(L):
            b = f"encoder.layer.{i}."
            for s_, t_ in (("q", "attention.self.query"), ("k", "attention.self.key"), ("v", "attention.self.value"),
                           ("o", "attention.output.dense"), ("fc1", "intermediate.dense"), ("fc2", "output.dense"),
                           ("ln1", "attention.output.LayerNorm"), ("ln2", "output.LayerNorm")):
                out[f"layer{i}.{s_}.weight"] = _np(sd[b + t_ + ".weight"])
                out[f"layer{i}.{s_}.bias"] = _np(sd[b + t_ + ".bias"])
        return out
    if cfg.arch == ARCH_WHISPER:
        pre = "encoder." if any(k.startswith("encoder.") for k in sd) else ""
        out["conv1.weight"] = _np(sd[pre + "conv1.weight"])
        out["conv1.bias"] = _np(sd[pre + "conv1.bias"])
        out["conv2.weight"] = _np(sd[pre + "conv2.weight"])
        out["conv2.bias"] = _np(sd[pre + "conv2.bias"])
        out["embed_positions"] = _np(sd[pre + "embed_positions.weight"])
        for i in range(L):
            b = f"{pre}layers.{i}."
            out[f"layer{i}.ln1.weight"] = _np(sd[b + "self_attn_layer_norm.weight"])
            out[f"layer{i}.ln1.bias"] = _np(sd[b + "self_attn_layer_norm.bias"])
            for s, t in (("q", "q_proj"), ("k", "k_proj"), ("v", "v_proj"), ("o", "out_proj")):
                out[f"layer{i}.{s}.weight"] = _np(sd[b + f"self_attn.{t}.weight"])
                if b + f"self_attn.{t}.bias" in sd:  # k_proj has no bias (modeling_whisper.py:279-282)
                    out[f"layer{i}.{s}.bias"] = _np(sd[b + f"self_attn.{t}.bias"])
            out[f"layer{i}.ln2.weight"] = _np(sd[b + "final_layer_norm.weight"])
            out[f"layer{i}.ln2.bias"] = _np(sd[b + "final_layer_norm.bias"])
            out[f"layer{i}.fc1.weight"] = _np(sd[b + "fc1.weight"])
            out[f"layer{i}.fc1.bias"] = _np(sd[b + "fc1.bias"])
            out[f"layer{i}.fc2.weight"] = _np(sd[b + "fc2.weight"])
            out[f"layer{i}.fc2.bias"] = _np(sd[b + "fc2.bias"])
        out["final_ln.weight"] = _np(sd[pre + "layer_norm.weight"])
        out["final_ln.bias"] = _np(sd[pre + "layer_norm.bias"])
        out["mel_filters"] = slaney_mel_filters(cfg.num_mel_bins)
        return out

    for i in range(7):
        b = f"feature_extractor.conv_layers.{i}."
        out[f"conv{i}.weight"] = _np(sd[b + "conv.weight"])
        if cfg.conv_bias:
            out[f"conv{i}.bias"] = _np(sd[b + "conv.bias"])
        if cfg.feat_extract_norm == "layer" or i == 0:   # 'group': only layer 0 carries a (Group)Norm
            out[f"conv{i}.ln.weight"] = _np(sd[b + "layer_norm.weight"])
            out[f"conv{i}.ln.bias"] = _np(sd[b + "layer_norm.bias"])
    if cfg.feat_proj_layer_norm:
        out["featproj.ln.weight"] = _np(sd["feature_projection.layer_norm.weight"])
        out["featproj.ln.bias"] = _np(sd["feature_projection.layer_norm.bias"])
    out["featproj.weight"] = _np(sd["feature_projection.projection.weight"])
    out["featproj.bias"] = _np(sd["feature_projection.projection.bias"])
    pc = "encoder.pos_conv_embed.conv."
    if pc + "parametrizations.weight.original0" in sd:
        w = fold_weight_norm(sd[pc + "parametrizations.weight.original0"], sd[pc + "parametrizations.weight.original1"])
    elif pc + "weight_g" in sd:
        w = fold_weight_norm(sd[pc + "weight_g"], sd[pc + "weight_v"])
    else:
        w = _np(sd[pc + "weight"])
    out["posconv.weight"] = w
    out["posconv.bias"] = _np(sd[pc + "bias"])
    for i in range(L):
        b = f"encoder.layers.{i}."
        out[f"layer{i}.ln1.weight"] = _np(sd[b + "layer_norm.weight"])
        out[f"layer{i}.ln1.bias"] = _np(sd[b + "layer_norm.bias"])
        for s, t in (("q", "q_proj"), ("k", "k_proj"), ("v", "v_proj"), ("o", "out_proj")):
            out[f"layer{i}.{s}.weight"] = _np(sd[b + f"attention.{t}.weight"])
            out[f"layer{i}.{s}.bias"] = _np(sd[b + f"attention.{t}.bias"])
        out[f"layer{i}.ln2.weight"] = _np(sd[b + "final_layer_norm.weight"])
        out[f"layer{i}.ln2.bias"] = _np(sd[b + "final_layer_norm.bias"])
        out[f"layer{i}.fc1.weight"] = _np(sd[b + "feed_forward.intermediate_dense.weight"])
        out[f"layer{i}.fc1.bias"] = _np(sd[b + "feed_forward.intermediate_dense.bias"])
        out[f"layer{i}.fc2.weight"] = _np(sd[b + "feed_forward.output_dense.weight"])
        out[f"layer{i}.fc2.bias"] = _np(sd[b + "feed_forward.output_dense.bias"])
        if cfg.family == "wavlm":
            out[f"layer{i}.gru.weight"] = _np(sd[b + "attention.gru_rel_pos_linear.weight"])
            out[f"layer{i}.gru.bias"] = _np(sd[b + "attention.gru_rel_pos_linear.bias"])
            out[f"layer{i}.gru.const"] = _np(sd[b + "attention.gru_rel_pos_const"]).reshape(-1)
    if cfg.family == "wavlm":
        out["rel_attn_embed"] = _np(sd["encoder.layers.0.attention.rel_attn_embed.weight"])
    out["final_ln.weight"] = _np(sd["encoder.layer_norm.weight"])
    out["final_ln.bias"] = _np(sd["encoder.layer_norm.bias"])
    return out


def random_init(cfg: EncoderConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random canonical weights of the right shapes (no network for checkpoints). Distributions follow the HF
    initialisers in spirit (normal(0, 0.02) linears, kaiming convs); biases and LayerNorm affines are
    randomised too, so that every parameter is exercised by parity tests and the benchmark."""
    rng = np.random.default_rng(seed)
    d, L, ffn, H = cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size, cfg.num_attention_heads
    hd = cfg.head_dim

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def ln(n):
        return (1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)).astype(np.float32), normal((n,), 0.05)

    out: Dict[str, np.ndarray] = {}
    for i in range(L):
        out[f"layer{i}.ln1.weight"], out[f"layer{i}.ln1.bias"] = ln(d)
        out[f"layer{i}.ln2.weight"], out[f"layer{i}.ln2.bias"] = ln(d)
        for s in ("q", "k", "v", "o"):
            out[f"layer{i}.{s}.weight"] = normal((d, d), 0.02)
            if not (cfg.arch == ARCH_WHISPER and s == "k"):
                out[f"layer{i}.{s}.bias"] = normal((d,), 0.02)
        out[f"layer{i}.fc1.weight"] = normal((ffn, d), 0.02)
        out[f"layer{i}.fc1.bias"] = normal((ffn,), 0.02)
        out[f"layer{i}.fc2.weight"] = normal((d, ffn), 0.02)
        out[f"layer{i}.fc2.bias"] = normal((d,), 0.02)
        if cfg.family == "wavlm":
            out[f"layer{i}.gru.weight"] = normal((8, hd), 0.05)
            out[f"layer{i}.gru.bias"] = normal((8,), 0.1)
            out[f"layer{i}.gru.const"] = (1.0 + 0.1 * rng.standard_normal(H, dtype=np.float32)).astype(np.float32)
    out["final_ln.weight"], out["final_ln.bias"] = ln(d)
    if cfg.arch == ARCH_TEXT:
        out["embed.word"] = normal((cfg.vocab_size, d), 0.5)
        out["embed.position"] = normal((cfg.max_position_embeddings, d), 0.5)
        out["embed.type"] = normal((cfg.type_vocab_size, d), 0.5)
        for k in ("word", "position"):   # nn.Embedding(padding_idx=pad): that row is zero in HF checkpoints
            out[f"embed.{k}"][cfg.pad_token_id] = 0.0
        return out
    if cfg.arch == ARCH_WHISPER:
        nm = cfg.num_mel_bins
        out["conv1.weight"] = normal((d, nm, 3), math.sqrt(2.0 / (nm * 3)))
        out["conv1.bias"] = normal((d,), 0.02)
        out["conv2.weight"] = normal((d, d, 3), math.sqrt(2.0 / (d * 3)))
        out["conv2.bias"] = normal((d,), 0.02)
        out["embed_positions"] = whisper_sinusoids(cfg.max_source_positions, d)
        out["mel_filters"] = slaney_mel_filters(nm)
        return out
    C = cfg.conv_dim[0]
    for i in range(7):
        cin = 1 if i == 0 else C
        k = cfg.conv_kernel[i]
        out[f"conv{i}.weight"] = normal((C, cin, k), math.sqrt(2.0 / (cin * k)))
        if cfg.conv_bias:
            out[f"conv{i}.bias"] = normal((C,), 0.05)
        if cfg.feat_extract_norm == "layer" or i == 0:
            out[f"conv{i}.ln.weight"], out[f"conv{i}.ln.bias"] = ln(C)
    if cfg.feat_proj_layer_norm:
        out["featproj.ln.weight"], out["featproj.ln.bias"] = ln(C)
    out["featproj.weight"] = normal((d, C), 1.0 / math.sqrt(C))
    out["featproj.bias"] = normal((d,), 0.02)
    kpos, g = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    out["posconv.weight"] = normal((d, d // g, kpos), 2.0 * math.sqrt(1.0 / (kpos * d)))
    out["posconv.bias"] = normal((d,), 0.02)
    if cfg.family == "wavlm":
        out["rel_attn_embed"] = normal((cfg.num_buckets, H), 1.0)
    return out


def load_checkpoint_dir(cfg: EncoderConfig, path: str) -> Dict[str, np.ndarray]:
    """Load an HF checkpoint directory (model.safetensors or pytorch_model.bin), a canonical .npz, or a single
    torch state-dict file (.pt / .bin / .pth — e.g. the LoRA classifier checkpoint of preprocess_speech_pretrained.py:173)."""
    npz = os.path.join(path, "serenc_weights.npz") if os.path.isdir(path) else path
    if npz.endswith(".npz") and os.path.exists(npz):
        with np.load(npz) as z:
            return {k: np.ascontiguousarray(z[k].astype(np.float32)) for k in z.files}
    import torch

    if os.path.isfile(path) and path.endswith((".pt", ".pth", ".bin")):
        return from_hf_state_dict(cfg, torch.load(path, map_location="cpu", weights_only=True))
    if os.path.isfile(path) and path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return from_hf_state_dict(cfg, load_file(path))
    st_path = os.path.join(path, "model.safetensors")
    bin_path = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(st_path):
        from safetensors.torch import load_file  # ships with transformers' dependency set

        return from_hf_state_dict(cfg, load_file(st_path))
    if os.path.exists(bin_path):
        return from_hf_state_dict(cfg, torch.load(bin_path, map_location="cpu", weights_only=True))
    raise OSError(f"no model.safetensors / pytorch_model.bin / serenc_weights.npz under '{path}'")


_NAME_RE = re.compile(r"^[a-z0-9_.]+$")


def validate_names(tensors: Mapping[str, np.ndarray]) -> None:
    for k in tensors:
        if not _NAME_RE.match(k):
            raise ValueError(f"bad tensor name {k!r}")
