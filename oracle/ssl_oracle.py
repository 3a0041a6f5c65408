"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product package.

CPU fp32 restatement (plain torch tensor ops on the CPU) of the arithmetic that the reference's
embedding-extraction scripts delegate to HuggingFace transformers:

    preprocessing/preprocess_speech.py:48-67    processor(...) ; model(**inputs, output_hidden_states=True)
    preprocessing/preprocess_whisper.py:48-76   processor(...)["input_features"] ; model.encoder(...)

The algorithm lives in a third-party dependency that is not vendored under /root/reference:
`transformers==4.47.1` (benchmark/requirements.txt:35); the copy installed in this image is 5.5.0 and
`HF:` citations below are relative to its `transformers/` directory. Every function names the HF lines it
follows. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / CPU baseline.

Pinning: the reference has no tests, golden vectors or fixtures for this path, so the
oracle is pinned against the reference implementation itself: oracle/make_golden.py imports HF transformers,
loads the same canonical weights into the HF modules, and (i) asserts oracle == HF to fp32 round-off and
(ii) writes HF's outputs to tests/golden/*.npz, which tests/test_oracle_golden.py re-checks without HF.

All functions take the canonical weight dict described in interspeech_ser_b200/weights.py (numpy fp32) and
process ONE utterance at a time, exactly as the reference scripts do (batch 1, no padding).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

W = Dict[str, np.ndarray]


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


# --------------------------------------------------------------------------------------------------
# feature extractor
# --------------------------------------------------------------------------------------------------
def normalize_waveform(x: np.ndarray) -> np.ndarray:
    """Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm, single utterance
    (HF models/wav2vec2/feature_extraction_wav2vec2.py:77-97): (x - mean) / sqrt(var + 1e-7)."""
    x = np.asarray(x, dtype=np.float32)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def w2v_num_frames(n: int, kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2)) -> int:
    """_get_feat_extract_output_lengths (HF models/wavlm/modeling_wavlm.py:640-659)."""
    for k, s in zip(kernels, strides):
        if n < k:
            return 0
        n = (n - k) // s + 1
    return n


# --------------------------------------------------------------------------------------------------
# wav2vec2 / HuBERT / WavLM
# --------------------------------------------------------------------------------------------------
def conv_feature_encoder(cfg, w: W, x: torch.Tensor) -> torch.Tensor:
    """feat_extract_norm='layer' (large / xlarge checkpoints): 7 x (Conv1d -> LayerNorm over channels -> exact GELU)
    (WavLMLayerNormConvLayer, HF modeling_wavlm.py:703-727; WavLMFeatureEncoder.forward :775-789).
    feat_extract_norm='group' (base checkpoints): layer 0 = Conv1d -> GroupNorm(groups = channels), i.e. per-channel
    statistics over TIME -> GELU (WavLMGroupNormConvLayer :730-751); layers 1-6 = Conv1d -> GELU
    (WavLMNoLayerNormConvLayer :682-700).  x: [L] normalised samples -> [T, 512]."""
    h = x[None, None, :]
    for i in range(7):
        b = _t(w[f"conv{i}.bias"]) if cfg.conv_bias else None
        h = F.conv1d(h, _t(w[f"conv{i}.weight"]), b, stride=cfg.conv_stride[i])
        if cfg.feat_extract_norm == "layer":
            h = h.transpose(-2, -1)
            h = F.layer_norm(h, (h.shape[-1],), _t(w[f"conv{i}.ln.weight"]), _t(w[f"conv{i}.ln.bias"]), 1e-5)
            h = h.transpose(-2, -1)
        elif i == 0:
            C = h.shape[1]
            h = F.group_norm(h, C, _t(w["conv0.ln.weight"]), _t(w["conv0.ln.bias"]), 1e-5)
        h = F.gelu(h)
    return h[0].transpose(0, 1).contiguous()


def feature_projection(cfg, w: W, feats: torch.Tensor) -> torch.Tensor:
    """LayerNorm(512) -> Linear(512 -> d) (WavLMFeatureProjection, HF modeling_wavlm.py:93-105); HuBERT-base skips
    the LayerNorm (`feat_proj_layer_norm=False`, HF models/hubert/modular_hubert.py:98-113)."""
    h = feats
    if getattr(cfg, "feat_proj_layer_norm", True):
        h = F.layer_norm(h, (h.shape[-1],), _t(w["featproj.ln.weight"]), _t(w["featproj.ln.bias"]), cfg.layer_norm_eps)
    return F.linear(h, _t(w["featproj.weight"]), _t(w["featproj.bias"]))


def pos_conv_embed(cfg, w: W, x: torch.Tensor) -> torch.Tensor:
    """Grouped Conv1d(k, padding=k//2) on the weight-norm-folded kernel, drop the last frame when k is even,
    exact GELU (WavLMPositionalConvEmbedding + WavLMSamePadLayer, HF modeling_wavlm.py:37-90). x: [T, d]."""
    k = cfg.num_conv_pos_embeddings
    h = x.transpose(0, 1)[None]
    h = F.conv1d(h, _t(w["posconv.weight"]), _t(w["posconv.bias"]), padding=k // 2, groups=cfg.num_conv_pos_embedding_groups)
    if k % 2 == 0:
        h = h[:, :, :-1]
    h = F.gelu(h)
    return h[0].transpose(0, 1)


def wavlm_bucket(rel: torch.Tensor, num_buckets: int = 320, max_distance: int = 800) -> torch.Tensor:
    """WavLMAttention._relative_positions_bucket (HF modeling_wavlm.py:253-271); rel = key_pos - query_pos."""
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = torch.log(rel.float() / max_exact)
    large = large / math.log(max_distance / max_exact)
    large = large * (nb - max_exact)
    large = (max_exact + large).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


def wavlm_position_bias(cfg, w: W, T: int) -> torch.Tensor:
    """compute_bias (HF modeling_wavlm.py:243-251): [H, T, T] from layer 0's rel_attn_embed."""
    ctx = torch.arange(T, dtype=torch.long)[:, None]
    mem = torch.arange(T, dtype=torch.long)[None, :]
    bucket = wavlm_bucket(mem - ctx, cfg.num_buckets, cfg.max_bucket_distance)
    return _t(w["rel_attn_embed"])[bucket].permute(2, 0, 1)


def self_attention(cfg, w: W, li: int, h: torch.Tensor, bias: torch.Tensor | None, whisper: bool = False,
                   n_keys: int | None = None) -> torch.Tensor:
    """Multi-head self-attention for one utterance (no padding, so no key mask; n_keys: the text encoder's padded
    sequences, where keys >= n_keys carry the -inf additive mask, HF RobertaSelfAttention modeling_roberta.py:190-254).
    wav2vec2/HuBERT: Wav2Vec2Attention (HF models/wav2vec2/modeling_wav2vec2.py:438-549);
    WavLM: torch MHA with the gated relative position bias as additive float mask (HF modeling_wavlm.py:147-228):
        gate_a, gate_b = sigmoid(gru_rel_pos_linear(x_head).view(2, 4).sum(-1));  g = a * (b * const - 1) + 2
        scores = (q / sqrt(dh)) k^T + g[:, None] * bias
    Whisper: q = (x Wq + bq) * dh^-0.5, k_proj without bias, no mask (HF modeling_whisper.py:279-357)."""
    T, d = h.shape
    H = cfg.num_attention_heads
    dh = d // H
    q = F.linear(h, _t(w[f"layer{li}.q.weight"]), _t(w[f"layer{li}.q.bias"]))
    kb = w.get(f"layer{li}.k.bias")
    k = F.linear(h, _t(w[f"layer{li}.k.weight"]), _t(kb) if kb is not None else None)
    v = F.linear(h, _t(w[f"layer{li}.v.weight"]), _t(w[f"layer{li}.v.bias"]))
    q = q.view(T, H, dh).transpose(0, 1) * (dh ** -0.5)
    k = k.view(T, H, dh).transpose(0, 1)
    v = v.view(T, H, dh).transpose(0, 1)
    scores = q @ k.transpose(1, 2)
    if bias is not None:
        xh = h.view(T, H, dh).transpose(0, 1)                                      # [H, T, dh]
        proj = F.linear(xh, _t(w[f"layer{li}.gru.weight"]), _t(w[f"layer{li}.gru.bias"]))  # [H, T, 8]
        proj = proj.view(H, T, 2, 4).sum(-1)
        gate = torch.sigmoid(proj)
        ga, gb = gate[..., 0], gate[..., 1]
        const = _t(w[f"layer{li}.gru.const"]).view(H, 1)
        g = ga * (gb * const - 1.0) + 2.0                                           # [H, T]
        scores = scores + g[:, :, None] * bias
    if n_keys is not None and n_keys < T:
        scores[:, :, n_keys:] = float("-inf")
    p = torch.softmax(scores, dim=-1)
    o = (p @ v).transpose(0, 1).reshape(T, d)
    return F.linear(o, _t(w[f"layer{li}.o.weight"]), _t(w[f"layer{li}.o.bias"]))


def encoder_layer(cfg, w: W, li: int, x: torch.Tensor, bias, whisper: bool = False) -> torch.Tensor:
    """Pre-LN block (WavLMEncoderLayerStableLayerNorm, HF modeling_wavlm.py:339-373;
    Wav2Vec2EncoderLayerStableLayerNorm, modeling_wav2vec2.py:612-655; WhisperEncoderLayer, modeling_whisper.py:361-414):
        x = x + Attn(LN1(x));  x = x + FC2(GELU(FC1(LN2(x))))."""
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), _t(w[f"layer{li}.ln1.weight"]), _t(w[f"layer{li}.ln1.bias"]), cfg.layer_norm_eps)
    x = x + self_attention(cfg, w, li, h, bias, whisper)
    h = F.layer_norm(x, (d,), _t(w[f"layer{li}.ln2.weight"]), _t(w[f"layer{li}.ln2.bias"]), cfg.layer_norm_eps)
    h = F.gelu(F.linear(h, _t(w[f"layer{li}.fc1.weight"]), _t(w[f"layer{li}.fc1.bias"])))
    return x + F.linear(h, _t(w[f"layer{li}.fc2.weight"]), _t(w[f"layer{li}.fc2.bias"]))


def encoder_layer_post_ln(cfg, w: W, li: int, x: torch.Tensor, bias, n_keys: int | None = None) -> torch.Tensor:
    """Post-LN block of the base-size checkpoints (WavLMEncoderLayer, HF modeling_wavlm.py:298-336; Wav2Vec2EncoderLayer,
    modeling_wav2vec2.py:576-609):  x = LN1(x + Attn(x));  x = LN2(x + FFN(x))."""
    d = x.shape[-1]
    x = x + self_attention(cfg, w, li, x, bias, n_keys=n_keys)
    x = F.layer_norm(x, (d,), _t(w[f"layer{li}.ln1.weight"]), _t(w[f"layer{li}.ln1.bias"]), cfg.layer_norm_eps)
    h = F.gelu(F.linear(x, _t(w[f"layer{li}.fc1.weight"]), _t(w[f"layer{li}.fc1.bias"])))
    x = x + F.linear(h, _t(w[f"layer{li}.fc2.weight"]), _t(w[f"layer{li}.fc2.bias"]))
    return F.layer_norm(x, (d,), _t(w[f"layer{li}.ln2.weight"]), _t(w[f"layer{li}.ln2.bias"]), cfg.layer_norm_eps)


def run_stack_post_ln(cfg, w: W, x: torch.Tensor, bias, n_keys: int | None = None) -> List[torch.Tensor]:
    """WavLMEncoder / Wav2Vec2Encoder (do_stable_layer_norm=False; HF modeling_wavlm.py:376-447): the encoder LayerNorm
    comes right after the positional conv, [0] = its output, [i] = output of layer i (already normalised)."""
    d = x.shape[-1]
    x = F.layer_norm(x, (d,), _t(w["final_ln.weight"]), _t(w["final_ln.bias"]), cfg.layer_norm_eps)
    hs = [x]
    for li in range(cfg.num_hidden_layers):
        x = encoder_layer_post_ln(cfg, w, li, x, bias, n_keys)
        hs.append(x)
    return hs


def run_stack(cfg, w: W, x: torch.Tensor, bias, whisper: bool = False) -> List[torch.Tensor]:
    """Encoder loop + final LayerNorm; returns the HF `hidden_states` tuple: L+1 tensors, [0] = stack input,
    [i] = residual stream after layer i, [L] = after the final LayerNorm
    (WavLMEncoderStableLayerNorm.forward, HF modeling_wavlm.py:450-522; WhisperEncoder.forward, modeling_whisper.py:593-647)."""
    hs = [x]
    L = cfg.num_hidden_layers
    for li in range(L):
        x = encoder_layer(cfg, w, li, x, bias, whisper)
        if li + 1 < L:
            hs.append(x)
    d = x.shape[-1]
    hs.append(F.layer_norm(x, (d,), _t(w["final_ln.weight"]), _t(w["final_ln.bias"]), cfg.layer_norm_eps))
    return hs


@torch.no_grad()
def w2v_hidden_states(cfg, w: W, wav: np.ndarray, normalize: bool = True) -> List[torch.Tensor]:
    """WavLMModel / Wav2Vec2Model / HubertModel forward with output_hidden_states=True for ONE utterance
    (HF modeling_wavlm.py:1039-1095). Returns L+1 tensors [T, d]."""
    x = _t(normalize_waveform(wav) if normalize else wav)
    feats = conv_feature_encoder(cfg, w, x)
    h = feature_projection(cfg, w, feats)
    h = h + pos_conv_embed(cfg, w, h)
    bias = wavlm_position_bias(cfg, w, h.shape[0]) if cfg.family == "wavlm" else None
    if not cfg.do_stable_layer_norm:
        return run_stack_post_ln(cfg, w, h, bias)
    return run_stack(cfg, w, h, bias)


# --------------------------------------------------------------------------------------------------
# Whisper
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def whisper_log_mel(w: W, wav: np.ndarray, n_fft: int = 400, hop: int = 160, n_samples: int = 480000) -> torch.Tensor:
    """WhisperFeatureExtractor: pad/truncate to 30 s, then _torch_extract_fbank_features
    (HF models/whisper/feature_extraction_whisper.py:135-164, :296-320). Returns [n_mels, 3000]."""
    x = np.zeros(n_samples, dtype=np.float32)
    n = min(len(wav), n_samples)
    x[:n] = np.asarray(wav[:n], dtype=np.float32)
    xt = torch.from_numpy(x)
    window = torch.hann_window(n_fft)
    stft = torch.stft(xt, n_fft, hop, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    mel = _t(w["mel_filters"]).T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


@torch.no_grad()
def whisper_hidden_states(cfg, w: W, mel: torch.Tensor) -> List[torch.Tensor]:
    """WhisperEncoder.forward(output_hidden_states=True) for one [n_mels, 3000] input
    (HF models/whisper/modeling_whisper.py:593-647): gelu(conv1) -> gelu(conv2, stride 2) -> + embed_positions -> stack."""
    h = mel[None]
    h = F.gelu(F.conv1d(h, _t(w["conv1.weight"]), _t(w["conv1.bias"]), padding=1))
    h = F.gelu(F.conv1d(h, _t(w["conv2.weight"]), _t(w["conv2.bias"]), stride=2, padding=1))
    h = h[0].transpose(0, 1) + _t(w["embed_positions"])
    return run_stack(cfg, w, h, None, whisper=True)


# --------------------------------------------------------------------------------------------------
# RoBERTa (text branch: preprocessing/preprocess_roberta.py:48-74)
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def roberta_hidden_states(cfg, w: W, input_ids: Sequence[int], n_valid: int | None = None) -> List[torch.Tensor]:
    """RobertaModel(input_ids, attention_mask, output_hidden_states=True) for ONE right-padded sequence
    (HF models/roberta/modeling_roberta.py: RobertaEmbeddings :56-159, RobertaSelfAttention :190-254, RobertaSelfOutput
    :334-345, RobertaIntermediate :377-389, RobertaOutput :392-403, RobertaLayer :406-468, RobertaEncoder :497-526).
    Embeddings: word[ids] + type[0] + position[pad + cumsum(ids != pad) * (ids != pad)] -> LayerNorm = hidden_states[0];
    every layer is post-LN:  x = LN(x + SelfOutput.dense(attn(x)));  x = LN(x + Output.dense(gelu(Intermediate.dense(x)))).
    Keys at pad positions get the -inf additive mask; ALL positions (pads included) are returned, as HF does.
    Returns L+1 tensors [T, d]."""
    ids = torch.as_tensor(list(input_ids), dtype=torch.long)
    pad = cfg.pad_token_id
    not_pad = ids.ne(pad)
    n_ids = int(not_pad.sum())
    if n_valid is None:
        n_valid = n_ids
    pos = torch.cumsum(not_pad.long(), 0) * not_pad.long() + pad
    x = _t(w["embed.word"])[ids] + _t(w["embed.type"])[0]
    x = x + _t(w["embed.position"])[pos]
    return run_stack_post_ln(cfg, w, x, None, n_keys=n_valid)


# --------------------------------------------------------------------------------------------------
# selection / pooling (the scripts' post-processing)
# --------------------------------------------------------------------------------------------------
def select_features(hidden_states: Sequence[torch.Tensor], layer: int = -1, average: bool = False) -> torch.Tensor:
    """preprocess_speech.py:56-67 / preprocess_whisper.py:61-73: hidden_states[layer] or mean of the last four."""
    if average:
        return torch.mean(torch.stack(list(hidden_states[-4:])), dim=0)
    return hidden_states[layer]


def masked_mean_pool(x: torch.Tensor, n_valid: int | None = None) -> torch.Tensor:
    """lora_wavlm/model.py:189-195 of the reference: sum over valid frames / number of valid frames."""
    n = x.shape[0] if n_valid is None else min(int(n_valid), x.shape[0])
    return x[:n].sum(dim=0) / float(n)


def whisper_keep_frames(n_samples: int, feat_dim: int) -> int:
    """preprocess_whisper.py:49-50,75: min(ceil(len/320), feats.shape[1]) — shape[1] is the hidden size (defect D2)."""
    return min(int(math.ceil(n_samples / 320)), feat_dim)
