"""Architecture constants of the encoders the reference extracts embeddings with.

The reference resolves these from the HF hub at run time (`AutoModel.from_pretrained(SSL_TYPE)`,
preprocessing/preprocess_speech.py:111-112); there is no network here, so the values of each checkpoint's
`config.json` that this path depends on are spelled out (SURVEY.md §8 "Model constants").
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, Tuple

ARCH_W2V = 0      # Wav2Vec2Model / HubertModel / WavLMModel
ARCH_WHISPER = 1  # WhisperModel.encoder
ARCH_TEXT = 2     # RobertaModel (preprocessing/preprocess_roberta.py of the reference)


@dataclass(frozen=True)
class EncoderConfig:
    name: str
    family: str                   # "wavlm" | "wav2vec2" | "hubert" | "whisper" | "roberta"
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    intermediate_size: int
    # wav2vec2-family feature encoder / positional conv
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_bias: bool = False
    feat_extract_norm: str = "layer"
    do_stable_layer_norm: bool = True
    feat_proj_layer_norm: bool = True
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    # WavLM
    num_buckets: int = 320
    max_bucket_distance: int = 800
    # Whisper
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    layer_norm_eps: float = 1e-5
    # feature-extractor behaviour (preprocessor_config.json)
    do_normalize: bool = True
    return_attention_mask: bool = True
    sampling_rate: int = 16000
    # text encoder (RobertaConfig)
    vocab_size: int = 0
    max_position_embeddings: int = 0
    type_vocab_size: int = 0
    pad_token_id: int = 1

    @property
    def arch(self) -> int:
        if self.family == "roberta":
            return ARCH_TEXT
        return ARCH_WHISPER if self.family == "whisper" else ARCH_W2V

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def to_dict(self) -> dict:
        return asdict(self)


_REGISTRY: Dict[str, EncoderConfig] = {}


def _reg(cfg: EncoderConfig, *aliases: str) -> EncoderConfig:
    for k in (cfg.name,) + aliases:
        _REGISTRY[k.lower()] = cfg
    return cfg


WAVLM_LARGE = _reg(
    EncoderConfig("microsoft/wavlm-large", "wavlm", 1024, 24, 16, 4096, conv_bias=False),
    "wavlm-large", "wavlm_large")
HUBERT_XLARGE = _reg(
    EncoderConfig("facebook/hubert-xlarge-ls960-ft", "hubert", 1280, 48, 16, 5120, conv_bias=True),
    "facebook/hubert-xlarge-ll60k", "hubert-xlarge-ls960-ft", "hubert-xlarge-ls960", "hubert-xlarge")
XLSR_2B = _reg(
    EncoderConfig("facebook/wav2vec2-xls-r-2b", "wav2vec2", 1920, 48, 16, 7680, conv_bias=True),
    "wav2vec2-xls-r-2b", "xls-r-2b")
WAV2VEC2_LARGE_LV60 = _reg(
    EncoderConfig("facebook/wav2vec2-large-lv60", "wav2vec2", 1024, 24, 16, 4096, conv_bias=True),
    "facebook/wav2vec2-large-robust", "wav2vec2-large-lv60", "wav2vec2-large-robust")
HUBERT_LARGE = _reg(
    EncoderConfig("facebook/hubert-large-ll60k", "hubert", 1024, 24, 16, 4096, conv_bias=True),
    "hubert-large-ll60k", "hubert-large")
# base-size checkpoints: GroupNorm feature encoder, post-LN transformer (benchmark/utils/etc.py:10-15 and
# configs/old/*wavlmbase* of the reference use them)
# preprocessor_config.json of wavlm-base / wavlm-base-plus: do_normalize false (only wavlm-large normalises), like the
# other GroupNorm checkpoints trained on un-normalised audio (recalled from the hub files, not re-verifiable offline)
WAVLM_BASE_PLUS = _reg(
    EncoderConfig("microsoft/wavlm-base-plus", "wavlm", 768, 12, 12, 3072, conv_bias=False, feat_extract_norm="group",
                  do_stable_layer_norm=False, do_normalize=False, return_attention_mask=True),
    "microsoft/wavlm-base", "wavlm-base-plus", "wavlm-base")
WAV2VEC2_BASE = _reg(
    EncoderConfig("facebook/wav2vec2-base", "wav2vec2", 768, 12, 12, 3072, conv_bias=False, feat_extract_norm="group",
                  do_stable_layer_norm=False, do_normalize=True, return_attention_mask=False),
    "facebook/wav2vec2-base-960h", "wav2vec2-base")
HUBERT_BASE = _reg(
    EncoderConfig("facebook/hubert-base-ls960", "hubert", 768, 12, 12, 3072, conv_bias=False, feat_extract_norm="group",
                  do_stable_layer_norm=False, feat_proj_layer_norm=False, do_normalize=False, return_attention_mask=False),
    "hubert-base-ls960", "hubert-base")
WHISPER_LARGE_V3 = _reg(
    EncoderConfig("openai/whisper-large-v3", "whisper", 1280, 32, 20, 5120, num_mel_bins=128),
    "whisper-large-v3")
WHISPER_LARGE_V2 = _reg(
    EncoderConfig("openai/whisper-large-v2", "whisper", 1280, 32, 20, 5120, num_mel_bins=80),
    "whisper-large-v2", "openai/whisper-large", "whisper-large")
WHISPER_MEDIUM = _reg(
    EncoderConfig("openai/whisper-medium", "whisper", 1024, 24, 16, 4096, num_mel_bins=80),
    "whisper-medium")

# Text encoders of the reference's text branch (preprocess_roberta.py:19 `--roberta_type`; post-LN BERT-style stacks)
_TEXT = dict(layer_norm_eps=1e-5, do_stable_layer_norm=False, vocab_size=50265, max_position_embeddings=514,
             type_vocab_size=1, pad_token_id=1)
ROBERTA_LARGE = _reg(EncoderConfig("roberta-large", "roberta", 1024, 24, 16, 4096, **_TEXT), "FacebookAI/roberta-large", "roberta")
ROBERTA_BASE = _reg(EncoderConfig("roberta-base", "roberta", 768, 12, 12, 3072, **_TEXT), "FacebookAI/roberta-base")
TINY_ROBERTA = _reg(EncoderConfig("tiny/roberta", "roberta", 128, 2, 2, 256, layer_norm_eps=1e-5, do_stable_layer_norm=False,
                                  vocab_size=300, max_position_embeddings=130, type_vocab_size=1, pad_token_id=1))

# Tiny configurations for fast known-answer tests (same code paths, seconds on CPU for the oracle).
TINY_WAVLM = _reg(
    EncoderConfig("tiny/wavlm", "wavlm", 128, 2, 2, 256, conv_bias=False,
                  num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4))
TINY_WAV2VEC2 = _reg(
    EncoderConfig("tiny/wav2vec2", "wav2vec2", 256, 2, 4, 512, conv_bias=True,
                  num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4))
TINY_HUBERT80 = _reg(  # head_dim 80 and a padded positional-conv group (like HuBERT-xlarge)
    EncoderConfig("tiny/hubert80", "hubert", 640, 2, 8, 1024, conv_bias=True,
                  num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=8))
TINY_W2V120 = _reg(    # head_dim 120 (like XLS-R-2b)
    EncoderConfig("tiny/w2v120", "wav2vec2", 1920, 1, 16, 768, conv_bias=True,
                  num_conv_pos_embeddings=15, num_conv_pos_embedding_groups=16))
TINY_WAVLM_BASE = _reg(   # GroupNorm conv encoder + post-LN layers + gated bias (like wavlm-base-plus)
    EncoderConfig("tiny/wavlm-base", "wavlm", 128, 2, 2, 256, conv_bias=False, feat_extract_norm="group",
                  do_stable_layer_norm=False, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4))
TINY_HUBERT_BASE = _reg(  # no feature-projection LayerNorm, conv bias (exercises the bias + GELU conv epilogue)
    EncoderConfig("tiny/hubert-base", "hubert", 256, 2, 4, 512, conv_bias=True, feat_extract_norm="group",
                  do_stable_layer_norm=False, feat_proj_layer_norm=False, num_conv_pos_embeddings=16,
                  num_conv_pos_embedding_groups=4))
TINY_WHISPER = _reg(
    EncoderConfig("tiny/whisper", "whisper", 128, 2, 2, 256, num_mel_bins=80))
TINY_WHISPER128 = _reg(
    EncoderConfig("tiny/whisper128", "whisper", 256, 1, 4, 512, num_mel_bins=128))


def get_config(name: str) -> EncoderConfig:
    key = name.lower().rstrip("/")
    if key in _REGISTRY:
        return _REGISTRY[key]
    base = key.split("/")[-1]
    if base in _REGISTRY:
        return _REGISTRY[base]
    # the reference raises OSError for an unknown model name (preprocess_speech.py:115-117)
    raise OSError(f"No architecture constants for '{name}'. Known: {sorted(set(c.name for c in _REGISTRY.values()))}")


def w2v_num_frames(n_samples: int, cfg: EncoderConfig = WAVLM_LARGE) -> int:
    """HF _get_feat_extract_output_lengths (modeling_wavlm.py:640-659)."""
    n = int(n_samples)
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        if n < k:
            return 0
        n = (n - k) // s + 1
    return n
