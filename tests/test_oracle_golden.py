"""The oracle (oracle/ssl_oracle.py) against golden vectors minted from the reference implementation
(HuggingFace transformers fp32 CPU forward on the same canonical weights; oracle/make_golden.py).
Weights are regenerated from the stored seed, never stored."""
import os

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import configs
from interspeech_ser_b200.weights import random_init, slaney_mel_filters
from oracle import ssl_oracle as O

WAVE_STD = 0.0886


def synth_wave(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * np.float32(WAVE_STD)).astype(np.float32)


def load(golden_dir, name):
    path = os.path.join(golden_dir, name.replace("/", "__") + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


@pytest.mark.parametrize("name", ["tiny/wavlm", "tiny/wav2vec2", "tiny/hubert80", "tiny/w2v120", "tiny/wavlm-base", "tiny/hubert-base"])
def test_w2v_oracle_matches_hf_golden(golden_dir, name):
    g = load(golden_dir, name)
    cfg = configs.get_config(name)
    w = random_init(cfg, int(g["seed"]))
    for j, n in enumerate(g["lengths"]):
        hs = O.w2v_hidden_states(cfg, w, synth_wave(int(g["wave_seed_base"]) + j, int(n)))
        assert len(hs) == cfg.num_hidden_layers + 1
        assert hs[0].shape == (O.w2v_num_frames(int(n)), cfg.hidden_size)
        pooled = np.stack([O.masked_mean_pool(h).numpy() for h in hs])
        np.testing.assert_allclose(pooled, g[f"pooled_{j}"], atol=2e-5, rtol=1e-4)
        np.testing.assert_allclose(hs[-1][:4].numpy(), g[f"last_{j}"], atol=5e-5, rtol=1e-4)
        ml4 = O.masked_mean_pool(O.select_features(hs, average=True)).numpy()
        np.testing.assert_allclose(ml4, g[f"meanlast4_pooled_{j}"], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("name", ["tiny/whisper", "tiny/whisper128"])
def test_whisper_oracle_matches_hf_golden(golden_dir, name):
    g = load(golden_dir, name)
    cfg = configs.get_config(name)
    w = random_init(cfg, int(g["seed"]))
    stride = int(g["mel_stride"])
    for j, n in enumerate(g["lengths"]):
        wav = synth_wave(int(g["wave_seed_base"]) + j, int(n))
        mel = O.whisper_log_mel(w, wav)
        assert mel.shape == (cfg.num_mel_bins, 3000)
        np.testing.assert_allclose(mel[:, ::stride].numpy(), g[f"mel_sub_{j}"], atol=1e-5)  # north_star tolerance is 1e-3
        hs = O.whisper_hidden_states(cfg, w, mel)
        keep = O.whisper_keep_frames(int(n), cfg.hidden_size)
        assert keep == int(g[f"keep_{j}"])
        pooled = np.stack([O.masked_mean_pool(h, keep).numpy() for h in hs])
        np.testing.assert_allclose(pooled, g[f"pooled_{j}"], atol=2e-5, rtol=1e-4)
        np.testing.assert_allclose(hs[-1][:4].numpy(), g[f"last_{j}"], atol=5e-5, rtol=1e-4)


def test_wavlm_large_oracle_matches_hf_golden(golden_dir):
    """Full-size WavLM-large (315 M parameters, random init seed 0), one 0.25 s utterance: seconds on CPU."""
    g = load(golden_dir, "microsoft/wavlm-large")
    cfg = configs.get_config("microsoft/wavlm-large")
    w = random_init(cfg, int(g["seed"]))
    j = 0
    hs = O.w2v_hidden_states(cfg, w, synth_wave(int(g["wave_seed_base"]) + j, int(g["lengths"][j])))
    pooled = np.stack([O.masked_mean_pool(h).numpy() for h in hs])
    assert pooled.shape == (25, 1024)
    scale = np.abs(g[f"pooled_{j}"]).max()
    assert np.abs(pooled - g[f"pooled_{j}"]).max() <= 2e-5 * max(1.0, scale)


def test_bucket_function_matches_hf_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "wavlm_buckets.npz"))
    mine = O.wavlm_bucket(torch.from_numpy(g["delta"]))
    assert np.array_equal(mine.numpy(), g["bucket"])
    assert len(np.unique(g["bucket"])) == 319  # 319 distinct buckets, saturation at |delta| >= 778
    assert g["bucket"][g["delta"] >= 778].min() == 319 and g["bucket"][g["delta"] <= -778].max() == 159


def test_logmel_oracle_on_structured_signals(golden_dir):
    from oracle.make_golden import logmel_signals
    g = np.load(os.path.join(golden_dir, "logmel_signals.npz"))
    w = {"mel_filters": slaney_mel_filters(128)}
    for name, x in logmel_signals().items():
        mel = O.whisper_log_mel(w, x)
        np.testing.assert_allclose(mel[:, ::25].numpy(), g[name], atol=2e-5, err_msg=name)
    # silence: everything sits at the clamp floor  (log10(1e-10) + 4) / 4 = -1.5
    assert np.allclose(g["silence_2s"], -1.5)


def test_normalize_waveform_properties():
    x = synth_wave(3, 16000) + 0.25
    y = O.normalize_waveform(x)
    assert abs(float(y.mean())) < 1e-5 and abs(float(y.var()) - 1.0) < 1e-3
    # idempotent up to the 1e-7 variance epsilon (relative 1e-7 / (2 var(x)) ~ 6e-6 here)
    np.testing.assert_allclose(O.normalize_waveform(y), y, rtol=5e-5, atol=1e-5)


def test_selection_and_pooling_semantics():
    hs = [torch.full((5, 3), float(i)) for i in range(6)]
    assert torch.equal(O.select_features(hs, layer=-1), hs[5])
    assert torch.equal(O.select_features(hs, layer=0), hs[0])
    assert torch.allclose(O.select_features(hs, average=True), torch.full((5, 3), 3.5))  # mean of states 2..5
    x = torch.arange(12, dtype=torch.float32).view(4, 3)
    assert torch.allclose(O.masked_mean_pool(x, 2), x[:2].mean(0))
    assert torch.allclose(O.masked_mean_pool(x), x.mean(0))
    # preprocess_whisper.py:75 (defect D2): cap is the hidden size, not 1500
    assert O.whisper_keep_frames(480000, 1280) == 1280 and O.whisper_keep_frames(64000, 1280) == 200
