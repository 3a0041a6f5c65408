"""Parity tests proper (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
(a) the CPU oracle on the same seeded inputs, (b) golden vectors minted from HuggingFace transformers, and
(c) size-independent properties at BASELINE sizes (batching invariance, permutation equivariance, determinism).

Tolerances are the ones BASELINE.json's north_star states: per-utterance pooled-embedding cosine >= 0.999 and
max relative error <= 2e-2 (bf16 operands, fp32 accumulate / residual stream); log-mel max abs error <= 1e-3 (fp32).
Max relative error is measured as max|a - b| / max|b| over the embedding vector.
"""
import os

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import configs
from interspeech_ser_b200.weights import random_init
from oracle import ssl_oracle as O

pytestmark = pytest.mark.gpu

COS_MIN = 0.999
REL_MAX = 2e-2
WAVE_STD = 0.0886


def synth_wave(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * np.float32(WAVE_STD)).astype(np.float32)


def check_embedding(got: torch.Tensor, ref: torch.Tensor, what: str):
    got, ref = got.float().cpu(), ref.float().cpu()
    cos = float(torch.nn.functional.cosine_similarity(got, ref, dim=0))
    rel = float((got - ref).abs().max() / ref.abs().max())
    assert cos >= COS_MIN and rel <= REL_MAX, f"{what}: cosine {cos:.6f}, max rel err {rel:.3e}"
    return cos, rel


_MODELS = {}


def get_model(name):
    if name not in _MODELS:
        from interspeech_ser_b200.modeling import SpeechEncoderModel, WhisperModel
        cfg = configs.get_config(name)
        w = random_init(cfg, 0)
        cls = WhisperModel if cfg.family == "whisper" else SpeechEncoderModel
        _MODELS[name] = (cfg, w, cls(cfg, w, 0))
    return _MODELS[name]


def load_golden(golden_dir, name):
    path = os.path.join(golden_dir, name.replace("/", "__") + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


# ------------------------------------------------------------------------------------------------
# wav2vec2 / HuBERT / WavLM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny/wavlm", "tiny/wav2vec2", "tiny/hubert80", "tiny/w2v120", "tiny/wavlm-base", "tiny/hubert-base"])
def test_w2v_every_hidden_state_vs_oracle_and_hf_golden(golden_dir, name):
    """Ragged batch incl. the edge lengths 400 / 401 / 719 / 720 (1, 1, 1, 2 frames): every hidden-state index, so
    both readings of the reference's layer-index defect are pinned."""
    cfg, w, model = get_model(name)
    g = load_golden(golden_dir, name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    L = cfg.num_hidden_layers
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    wav = torch.from_numpy(np.concatenate(waves)).cuda()
    frames, pooled, offs, idx = model.engine.encode_w2v(wav, starts, lens, normalize=True, layers=range(L + 1), want_frames=True, want_pooled=True)
    torch.cuda.synchronize()
    assert idx == list(range(L + 1)) and offs[-1] == sum(O.w2v_num_frames(n) for n in lens)
    frames, pooled = frames.cpu(), pooled.cpu()
    for b, wv in enumerate(waves):
        hs = O.w2v_hidden_states(cfg, w, wv)
        assert offs[b + 1] - offs[b] == hs[0].shape[0]
        for i in range(L + 1):
            check_embedding(pooled[i, b], O.masked_mean_pool(hs[i]), f"{name} utt{b} hs{i} vs oracle")
            check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"{name} utt{b} hs{i} vs HF golden")
            f = frames[i, offs[b]:offs[b + 1]]
            assert float((f - hs[i]).abs().max() / hs[i].abs().max()) <= 4e-2      # frame level, looser than pooled
            assert torch.allclose(f.mean(0), pooled[i, b], atol=1e-5, rtol=1e-5)    # pooling == masked mean of the frames


def test_wavlm_large_vs_hf_golden(golden_dir):
    """Full-size WavLM-large (BASELINE metric model), 0.25 s / 4 s / 12 s utterances in one packed batch."""
    name = "microsoft/wavlm-large"
    cfg, w, model = get_model(name)
    g = load_golden(golden_dir, name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    L = cfg.num_hidden_layers
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    wav = torch.from_numpy(np.concatenate(waves)).cuda()
    _, pooled, offs, _ = model.engine.encode_w2v(wav, starts, lens, normalize=True, layers=range(L + 1), want_frames=False, want_pooled=True)
    torch.cuda.synchronize()
    worst = (1.0, 0.0)
    for b in range(len(lens)):
        for i in range(L + 1):
            cos, rel = check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"wavlm-large utt{b} hs{i}")
            worst = (min(worst[0], cos), max(worst[1], rel))
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True)
    for b in range(len(lens)):
        check_embedding(res.pooled[b], torch.from_numpy(g[f"meanlast4_pooled_{b}"]), f"wavlm-large utt{b} mean-last-4")
    print(f"wavlm-large worst cosine {worst[0]:.6f}, worst max-rel {worst[1]:.3e}")


@pytest.mark.parametrize("name", ["facebook/hubert-xlarge-ls960-ft", "facebook/wav2vec2-xls-r-2b"])
def test_xlarge_models_vs_hf_golden(golden_dir, name):
    """Full-size HuBERT-xlarge (48 layers, d = 1280, head_dim 80) and XLS-R-2b (48 layers, d = 1920, head_dim 120):
    BASELINE configs[2] / [3] architectures, every hidden state of a 0.25 s + 6 s packed batch against the HF fp32
    forward on the same seeded weights (the wide-head tcgen05 attention, the 10 / 15-vector LayerNorm widths and the
    padded positional-conv groups at their real sizes)."""
    g = load_golden(golden_dir, name)
    cfg, w, model = get_model(name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    L = cfg.num_hidden_layers
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    wav = torch.from_numpy(np.concatenate(waves)).cuda()
    _, pooled, offs, _ = model.engine.encode_w2v(wav, starts, lens, normalize=True, layers=range(L + 1), want_frames=False, want_pooled=True)
    torch.cuda.synchronize()
    worst = (1.0, 0.0)
    for b in range(len(lens)):
        for i in range(L + 1):
            cos, rel = check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"{name} utt{b} hs{i}")
            worst = (min(worst[0], cos), max(worst[1], rel))
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True)
    for b in range(len(lens)):
        check_embedding(res.pooled[b], torch.from_numpy(g[f"meanlast4_pooled_{b}"]), f"{name} utt{b} mean-last-4")
    print(f"{name} worst cosine {worst[0]:.6f}, worst max-rel {worst[1]:.3e}")
    _MODELS.pop(name, None)  # free the host copy of the weights (3.8 / 8.7 GB)
    del model


def test_baseline_config2_length_bucketed_batching_invariance():
    """BASELINE configs[2]: HuBERT-xlarge over variable-length 2-12 s utterances. Size-independent property at full
    model size: an utterance's embedding in a ragged packed batch (scheduler order) equals, bit for bit, the one it
    gets alone or in a permuted batch - packed rows, ragged attention tiles and per-utterance pooling never mix
    utterances, and no kernel's reduction order depends on the batch."""
    from interspeech_ser_b200 import scheduler
    name = "facebook/hubert-xlarge-ls960-ft"
    cfg, w, model = get_model(name)
    rng = np.random.default_rng(11)
    lens = [int(v) for v in rng.integers(2 * 16000, 12 * 16000, size=12)]
    waves = [synth_wave(700 + j, n) for j, n in enumerate(lens)]
    batches = scheduler.make_batches(cfg, lens, frame_budget=4096)
    assert len(batches) >= 1 and sorted(i for b in batches for i in b.indices) == list(range(12))
    got = {}
    for bt in batches:
        res = model.extract([waves[i] for i in bt.indices], average=True, want_frames=False, want_pooled=True).pooled.cpu()
        for j, i in enumerate(bt.indices):
            got[i] = res[j]
    for i in (0, 5, 11):
        alone = model.extract([waves[i]], average=True, want_frames=False, want_pooled=True).pooled.cpu()[0]
        assert torch.equal(alone, got[i])
    rev = model.extract([waves[i] for i in reversed(range(12))], average=True, want_frames=False, want_pooled=True).pooled.cpu()
    for j, i in enumerate(reversed(range(12))):
        assert torch.equal(rev[j], got[i])
    assert all(torch.isfinite(v).all() for v in got.values())
    _MODELS.pop(name, None)


def test_baseline_config0_batching_invariance_and_determinism():
    """BASELINE configs[0]: WavLM-large, batch 8 x 4 s. The embedding of an utterance must not depend on its batch
    (the reference runs batch 1): packed batch == one-by-one == permuted batch, bit for bit, and twice the same."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    waves = [synth_wave(100 + j, 64000) for j in range(8)]
    a = model.extract(waves, average=True, want_frames=False, want_pooled=True).pooled.cpu()
    b = model.extract(waves, average=True, want_frames=False, want_pooled=True).pooled.cpu()
    assert torch.equal(a, b)
    perm = [3, 0, 7, 1, 6, 2, 5, 4]
    c = model.extract([waves[i] for i in perm], average=True, want_frames=False, want_pooled=True).pooled.cpu()
    assert torch.equal(c, a[perm])
    for j in (0, 5):
        one = model.extract([waves[j]], average=True, want_frames=False, want_pooled=True).pooled.cpu()
        assert torch.equal(one[0], a[j])
    assert a.shape == (8, 1024) and torch.isfinite(a).all()


def test_upload_ring_overlapped_batches_match_direct_calls():
    """extract_pinned (two-slot upload ring on a copy stream) over a run of different batches, sizes growing
    and shrinking so that slots are re-allocated and reused, without any host synchronisation in between: every result
    equals the plain device-resident call bit for bit."""
    cfg, w, model = get_model("tiny/wavlm")
    batches = []
    for k, (nb, n) in enumerate([(3, 8000), (5, 16000), (2, 4001), (6, 16000), (1, 32000), (4, 8000), (3, 8000)]):
        lens = [n - 37 * j for j in range(nb)]
        host = torch.from_numpy(np.concatenate([synth_wave(500 + 10 * k + j, m) for j, m in enumerate(lens)])).pin_memory()
        batches.append((host, lens))
    got = [model.extract_pinned(h, lens, average=True, want_frames=False, want_pooled=True).pooled for h, lens in batches]
    torch.cuda.synchronize()
    for (h, lens), g in zip(batches, got):
        ref = model.extract_device(h.to(model.device), lens, average=True, want_frames=False, want_pooled=True).pooled
        assert torch.equal(g, ref)


def test_encode_call_is_cuda_graph_capturable():
    """SURVEY 8b ownership row: no hidden allocation or synchronisation on the call path, so one encode call through
    the C ABI (span upload, every kernel, pooling) can be captured into a CUDA graph and replayed on new waveform
    contents in the same buffers; the replay must equal the eager call bit for bit."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    lens = [64000] * 8
    wa = torch.from_numpy(np.concatenate([synth_wave(200 + j, 64000) for j in range(8)]))
    wb = torch.from_numpy(np.concatenate([synth_wave(300 + j, 64000) for j in range(8)]))
    static = torch.empty(8 * 64000, dtype=torch.float32, device=model.device)
    static.copy_(wa)
    eager_a = model.extract_device(static, lens, average=True, use_graph=False).pooled.clone()   # also sizes the workspace outside the capture
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = model.extract_device(static, lens, average=True).pooled
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager_a)
    static.copy_(wb)
    g.replay()
    torch.cuda.synchronize()
    got_b = out.clone()
    eager_b = model.extract_device(static, lens, average=True, use_graph=False).pooled
    assert torch.equal(got_b, eager_b) and not torch.equal(got_b, eager_a)


def test_lora_checkpoint_merged_at_load(tmp_path):
    """preprocess_speech_pretrained.py:120-178: a peft LoRA (r=8, alpha=16, q_proj/v_proj) classifier state dict,
    loaded from its .pt file, must give the embeddings of base + adapter (oracle run on independently merged weights)."""
    from interspeech_ser_b200.modeling import AutoModel
    from interspeech_ser_b200.weights import from_hf_state_dict
    cfg = configs.get_config("tiny/wavlm")
    w = random_init(cfg, 3)
    rng = np.random.default_rng(5)
    hf_names = {"q": "q_proj", "k": "k_proj", "v": "v_proj", "o": "out_proj"}
    merged = {k: v.copy() for k, v in w.items()}
    sd = {}
    # hand-built peft layout around the canonical -> HF name map (independent of weights.merge_lora)
    from oracle.make_golden import hf_model
    for k, v in hf_model(cfg, w).state_dict().items():
        mod, _, leaf = k.rpartition(".")
        if mod.endswith(("q_proj", "v_proj")):
            sd[f"wavlm.base_model.model.{mod}.base_layer.{leaf}"] = v
        else:
            sd[f"wavlm.base_model.model.{k}"] = v
    for i in range(cfg.num_hidden_layers):
        for s_ in ("q", "v"):
            a = (rng.standard_normal((8, cfg.hidden_size)) * 0.2).astype(np.float32)
            b = (rng.standard_normal((cfg.hidden_size, 8)) * 0.2).astype(np.float32)
            mod = f"wavlm.base_model.model.encoder.layers.{i}.attention.{hf_names[s_]}"
            sd[mod + ".lora_A.default.weight"] = torch.from_numpy(a)
            sd[mod + ".lora_B.default.weight"] = torch.from_numpy(b)
            merged[f"layer{i}.{s_}.weight"] = (w[f"layer{i}.{s_}.weight"].astype(np.float64) + 2.0 * (b.astype(np.float64) @ a.astype(np.float64))).astype(np.float32)
    sd["classifier.0.weight"] = torch.zeros(512, cfg.hidden_size)
    path = str(tmp_path / "whisper_lora_ser.pt")
    torch.save(sd, path)
    model = AutoModel.from_pretrained(path, device=0, config_name="tiny/wavlm")
    waves = [synth_wave(900 + j, n) for j, n in enumerate((16000, 5000))]
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True)
    for b_, wv in enumerate(waves):
        hs = O.w2v_hidden_states(cfg, merged, wv)
        ref = torch.stack(hs[-4:]).mean(0)
        check_embedding(res.pooled[b_], O.masked_mean_pool(ref), f"lora utt{b_}")
        hs0 = O.w2v_hidden_states(cfg, w, wv)      # the adapter must matter, or this test proves nothing
        assert float((torch.stack(hs0[-4:]).mean(0) - ref).abs().max()) > 1e-2
    assert np.allclose(from_hf_state_dict(cfg, sd)["layer1.v.weight"], merged["layer1.v.weight"], atol=1e-6)


def test_hf_call_surface_w2v():
    """processor(...) -> model(**inputs, output_hidden_states=True) exactly as preprocess_speech.py:48-67 uses them."""
    from interspeech_ser_b200.modeling import AutoFeatureExtractor
    cfg, w, model = get_model("tiny/wavlm")
    proc = AutoFeatureExtractor.from_pretrained("tiny/wavlm", model=model)
    y = synth_wave(3, 17777)
    inputs = proc(y, sampling_rate=16000, return_tensors="pt", padding=True)
    assert set(inputs) == {"input_values", "attention_mask"}
    np.testing.assert_allclose(inputs["input_values"][0].cpu().numpy(), O.normalize_waveform(y), atol=2e-5)
    inputs = {k: v.to("cuda") for k, v in inputs.items()}
    inputs["output_hidden_states"] = True
    out = model.eval().to("cuda")(**inputs)
    hs = O.w2v_hidden_states(cfg, w, y)
    assert len(out.hidden_states) == len(out["hidden_states"]) == cfg.num_hidden_layers + 1
    assert out.hidden_states[0].shape == (1, 55, cfg.hidden_size) and out.last_hidden_state is out.hidden_states[-1]
    for i, h in enumerate(hs):
        check_embedding(out["hidden_states"][i][0].mean(0), h.mean(0), f"hf-surface hs{i}")
    feats = torch.mean(torch.stack(out.hidden_states[-4:]), dim=0).squeeze(0)       # the script's own post-processing
    check_embedding(feats.mean(0), O.select_features(hs, average=True).mean(0), "hf-surface mean-last-4")
    with pytest.raises(ValueError):
        proc(y, sampling_rate=8000)
    # padded batch with attention_mask (float mask as the benchmark caller passes, train_cat_ser.py:173-175)
    y2 = synth_wave(4, 4001)
    batch = proc([y, y2], sampling_rate=16000, return_tensors="pt", padding=True)
    out2 = model(batch["input_values"], attention_mask=batch["attention_mask"].float())
    assert out2.last_hidden_state.shape == (2, 55, cfg.hidden_size)
    t2 = O.w2v_num_frames(4001)
    assert torch.count_nonzero(out2.last_hidden_state[1, t2:]) == 0
    check_embedding(out2.last_hidden_state[1, :t2].mean(0), O.w2v_hidden_states(cfg, w, y2)[-1].mean(0), "padded batch utt1")
    with pytest.raises(ValueError):
        model.extract([np.zeros(399, dtype=np.float32)])          # shorter than the receptive field


# ------------------------------------------------------------------------------------------------
# Whisper
# ------------------------------------------------------------------------------------------------
def test_logmel_structured_signals(golden_dir):
    """silence, impulse, 440 Hz / 7999 Hz sines, noise; 1 s / 30 s / 31 s. fp32 tolerance 1e-3."""
    from oracle.make_golden import logmel_signals
    cfg, w, model = get_model("tiny/whisper128")
    g = np.load(os.path.join(golden_dir, "logmel_signals.npz"))
    sigs = logmel_signals()
    names = list(sigs)
    waves = [sigs[n] for n in names]
    lens = [len(x) for x in waves]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    mel = model.engine.logmel(torch.from_numpy(np.concatenate(waves)).cuda(), starts, lens).cpu()
    for b, n in enumerate(names):
        ref = O.whisper_log_mel(w, waves[b])
        assert float((mel[b] - ref).abs().max()) <= 1e-3, n
        assert float((mel[b][:, ::25] - torch.from_numpy(g[n])).abs().max()) <= 1e-3, n      # HF golden
    assert torch.all(mel[names.index("silence_2s")] == -1.5)


@pytest.mark.parametrize("name", ["tiny/whisper", "tiny/whisper128"])
def test_whisper_vs_oracle_and_hf_golden(golden_dir, name):
    cfg, w, model = get_model(name)
    g = load_golden(golden_dir, name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    L = cfg.num_hidden_layers
    clipped = [wv[:480000] for wv in waves]
    cl = [len(x) for x in clipped]
    starts = np.concatenate([[0], np.cumsum(cl)[:-1]]).tolist()
    mel = model.engine.logmel(torch.from_numpy(np.concatenate(clipped)).cuda(), starts, cl)
    keep = [O.whisper_keep_frames(n, cfg.hidden_size) for n in lens]
    frames, pooled, idx = model.engine.encode_whisper(mel, layers=range(L + 1), n_keep=keep, want_frames=True, want_pooled=True)
    torch.cuda.synchronize()
    stride = int(g["mel_stride"])
    for b in range(len(lens)):
        assert float((mel[b].cpu()[:, ::stride] - torch.from_numpy(g[f"mel_sub_{b}"])).abs().max()) <= 1e-3
        assert keep[b] == int(g[f"keep_{b}"])
        hs = O.whisper_hidden_states(cfg, w, O.whisper_log_mel(w, waves[b]))
        for i in range(L + 1):
            check_embedding(pooled[i, b], O.masked_mean_pool(hs[i], keep[b]), f"{name} utt{b} hs{i} vs oracle")
            check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"{name} utt{b} hs{i} vs HF golden")


def test_whisper_large_v3_vs_hf_golden(golden_dir):
    name = "openai/whisper-large-v3"
    g = load_golden(golden_dir, name)
    cfg, w, model = get_model(name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    L = cfg.num_hidden_layers
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    mel = model.engine.logmel(torch.from_numpy(np.concatenate(waves)).cuda(), starts, lens)
    keep = [int(g[f"keep_{b}"]) for b in range(len(lens))]
    _, pooled, _ = model.engine.encode_whisper(mel, layers=range(L + 1), n_keep=keep, want_frames=False, want_pooled=True)
    for b in range(len(lens)):
        assert float((mel[b].cpu()[:, ::int(g["mel_stride"])] - torch.from_numpy(g[f"mel_sub_{b}"])).abs().max()) <= 1e-3
        for i in range(L + 1):
            check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"whisper-large-v3 utt{b} hs{i}")
    _MODELS.pop(name, None)  # free 1.3 GB of host weights


def test_hf_call_surface_whisper():
    from interspeech_ser_b200.modeling import AutoProcessor
    cfg, w, model = get_model("tiny/whisper")
    proc = AutoProcessor.from_pretrained("tiny/whisper", model=model)
    y = synth_wave(9, 80000)
    feats = proc(y, sampling_rate=16000, return_tensors="pt")["input_features"].to("cuda")
    assert feats.shape == (1, cfg.num_mel_bins, 3000)
    out = model.encoder(feats, output_hidden_states=True)
    assert len(out.hidden_states) == cfg.num_hidden_layers + 1 and out["hidden_states"][-1].shape == (1, 1500, cfg.hidden_size)
    hs = O.whisper_hidden_states(cfg, w, O.whisper_log_mel(w, y))
    for i, h in enumerate(hs):
        check_embedding(out.hidden_states[i][0].mean(0), h.mean(0), f"whisper hf-surface hs{i}")
    with pytest.raises(ValueError):
        model.encoder(feats[:, :, :2999].contiguous())       # HF raises for != 3000 mel frames (modeling_whisper.py:613-617)


# ------------------------------------------------------------------------------------------------
# CLI contract (file naming, [T, D] fp32 on disk, D1/D2 behaviours, error-and-continue)
# ------------------------------------------------------------------------------------------------
def test_cli_speech_contract(tmp_path, capsys):
    from interspeech_ser_b200 import audio_io
    from interspeech_ser_b200.cli import main_speech
    wav_dir, out_dir = tmp_path / "wav", tmp_path / "feat"
    wav_dir.mkdir()
    lens = {"MSP-PODCAST_0001_0001.wav": 16000, "MSP-PODCAST_0001_0002.wav": 40001, "MSP-PODCAST_0002_0001.wav": 4001}
    for k, n in lens.items():
        audio_io.write_wav(str(wav_dir / k), synth_wave(hash(k) % 1000, n) * 3)
    audio_io.write_wav(str(wav_dir / "too_short.wav"), synth_wave(1, 200))
    (wav_dir / "broken.wav").write_bytes(b"garbage")
    rc = main_speech(["--ssl_type", "tiny/wavlm", "--wav_dir", str(wav_dir), "--save_path", str(out_dir), "--random_init",
                      "--use_average", "y", "--pooled_path", str(tmp_path / "pooled.pt")])
    assert rc == 0
    text = capsys.readouterr().out
    assert text.count("Failed to process") == 2 and "too_short.wav" in text and "broken.wav" in text
    cfg, w, _ = get_model("tiny/wavlm")
    for k, n in lens.items():
        t = torch.load(out_dir / (os.path.splitext(k)[0] + ".pt"))
        assert t.dtype == torch.float32 and t.device.type == "cpu" and t.is_contiguous()
        assert t.shape == (O.w2v_num_frames(n), cfg.hidden_size)
        y, _ = audio_io.load_audio(str(wav_dir / k))
        ref = O.select_features(O.w2v_hidden_states(cfg, w, y), average=True)
        check_embedding(t.mean(0), ref.mean(0), f"cli {k}")
    assert not (out_dir / "too_short.pt").exists()
    pooled = torch.load(tmp_path / "pooled.pt")
    assert pooled["embeddings"].shape == (3, cfg.hidden_size) and len(pooled["names"]) == 3
    # --n_layer is honoured (intended behaviour); the literal directory-count indexing is opt-in (defect D1)
    out2 = tmp_path / "feat_l0"
    # (two files per decode window: the streamed driver loop with its one-window-ahead decode)
    assert main_speech(["--ssl_type", "tiny/wavlm", "--wav_dir", str(wav_dir), "--save_path", str(out2), "--random_init", "--n_layer", "0",
                        "--window_files", "2"]) == 0
    assert sorted(os.listdir(out2)) == sorted(os.path.splitext(k)[0] + ".pt" for k in lens)
    k = "MSP-PODCAST_0001_0001.wav"
    y, _ = audio_io.load_audio(str(wav_dir / k))
    t0 = torch.load(out2 / "MSP-PODCAST_0001_0001.pt")
    check_embedding(t0.mean(0), O.w2v_hidden_states(cfg, w, y)[0].mean(0), "cli n_layer 0")
    out3 = tmp_path / "feat_compat"
    assert main_speech(["--ssl_type", "tiny/wavlm", "--wav_dir", str(wav_dir), "--save_path", str(out3), "--random_init",
                        "--compat_layer_from_dir_count"]) == 0
    assert torch.equal(torch.load(out3 / "MSP-PODCAST_0001_0001.pt"), t0)     # empty dir -> hidden_states[0]
    capsys.readouterr()
    assert main_speech(["--ssl_type", "no/such-model", "--wav_dir", str(wav_dir), "--save_path", str(out3)]) == 1
    assert "No pretrained model found with the name no/such-model" in capsys.readouterr().out


def test_cli_whisper_contract(tmp_path, capsys):
    from interspeech_ser_b200 import audio_io
    from interspeech_ser_b200.cli import main_whisper
    wav_dir, out_dir = tmp_path / "wav", tmp_path / "feat"
    wav_dir.mkdir()
    cfg, w, _ = get_model("tiny/whisper")
    lens = {"a.wav": 16000, "b.wav": 100000}
    for k, n in lens.items():
        audio_io.write_wav(str(wav_dir / k), synth_wave(len(k) + n % 7, n) * 3)
    assert main_whisper(["--ssl_type", "tiny/whisper", "--wav_dir", str(wav_dir), "--save_path", str(out_dir), "--random_init"]) == 0
    for k, n in lens.items():
        t = torch.load(out_dir / (os.path.splitext(k)[0] + ".pt"))
        keep = O.whisper_keep_frames(n, cfg.hidden_size)            # literal: min(ceil(n/320), hidden_size=128)
        assert t.shape == (keep, cfg.hidden_size) and t.dtype == torch.float32 and t.is_contiguous()
        assert t.untyped_storage().nbytes() == keep * cfg.hidden_size * 4   # no 1500-frame storage behind a view (D3)
        y, _ = audio_io.load_audio(str(wav_dir / k))
        hs = O.whisper_hidden_states(cfg, w, O.whisper_log_mel(w, y))
        check_embedding(t.mean(0), hs[-1][:keep].mean(0), f"whisper cli {k}")
    out2 = tmp_path / "feat1500"
    assert main_whisper(["--ssl_type", "tiny/whisper", "--wav_dir", str(wav_dir), "--save_path", str(out2), "--random_init", "--crop_cap_1500"]) == 0
    assert torch.load(out2 / "b.pt").shape == (313, cfg.hidden_size)
