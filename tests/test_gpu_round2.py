"""Round-2 parity tests (`-m gpu`, all through the C ABI): the benched sizes, the long and outlier goldens, the
int16 / extract_features / weighted-sum / request-batching / CUDA-graph surfaces, and the text branch (RoBERTa).

Tolerances as in test_gpu_models.py (north_star): pooled cosine >= 0.999, max|a - b| / max|b| <= 2e-2."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import configs
from interspeech_ser_b200.weights import random_init
from oracle import ssl_oracle as O
from test_gpu_models import check_embedding, get_model, load_golden, synth_wave, _MODELS

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pooled_all_states(model, waves):
    lens = [len(w) for w in waves]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    wav = torch.from_numpy(np.concatenate(waves)).cuda()
    L = model.cfg.num_hidden_layers
    _, pooled, offs, _ = model.engine.encode_w2v(wav, starts, lens, normalize=True, layers=range(L + 1), want_frames=False, want_pooled=True)
    torch.cuda.synchronize()
    return pooled, offs


# ------------------------------------------------------------------------------------------------
# goldens at the benched lengths and with trained-checkpoint statistics
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["microsoft/wavlm-large", "facebook/hubert-xlarge-ls960-ft", "facebook/wav2vec2-xls-r-2b"])
def test_long_utterance_vs_hf_golden(golden_dir, name):
    """20 s (T = 999: BASELINE configs[4]'s longest utterance, eight query tiles and the full-size WavLM bias window) for
    WavLM-large, 8 s (T = 399, configs[3]) for the wide-head models: every hidden state against the HF fp32 forward."""
    g = load_golden(golden_dir, name + "__long")
    cfg, w, model = get_model(name)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    pooled, offs = pooled_all_states(model, waves)
    assert offs[-1] == sum(O.w2v_num_frames(n) for n in lens)
    worst = (1.0, 0.0)
    for b in range(len(lens)):
        for i in range(cfg.num_hidden_layers + 1):
            cos, rel = check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"{name} long utt{b} hs{i}")
            worst = (min(worst[0], cos), max(worst[1], rel))
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True)
    for b in range(len(lens)):
        check_embedding(res.pooled[b], torch.from_numpy(g[f"meanlast4_pooled_{b}"]), f"{name} long utt{b} mean-last-4")
    print(f"{name} long: worst cosine {worst[0]:.6f}, worst max-rel {worst[1]:.3e}")
    if name != "microsoft/wavlm-large":
        _MODELS.pop(name, None)


def test_outlier_channels_vs_hf_golden(golden_dir):
    """The only offline proxy for a trained checkpoint's statistics: four residual-stream channels ~300x above the rest
    (|x| up to 900 against ~3) and one LayerNorm gain x20. bf16 GEMM operands with an fp32 residual stream and fp32
    LayerNorm statistics have to stay inside the same tolerance as the N(0, 0.02) weights."""
    from interspeech_ser_b200.modeling import SpeechEncoderModel
    from oracle.make_golden import outlier_init

    g = load_golden(golden_dir, "microsoft/wavlm-large__outlier")
    cfg = configs.get_config("microsoft/wavlm-large")
    model = SpeechEncoderModel(cfg, outlier_init(cfg, int(g["seed"])), 0)
    lens = [int(n) for n in g["lengths"]]
    waves = [synth_wave(int(g["wave_seed_base"]) + j, n) for j, n in enumerate(lens)]
    pooled, _ = pooled_all_states(model, waves)
    worst = (1.0, 0.0)
    for b in range(len(lens)):
        assert float(np.abs(g[f"pooled_{b}"]).max()) > 100.0          # the fixture really has outliers
        for i in range(cfg.num_hidden_layers + 1):
            cos, rel = check_embedding(pooled[i, b], torch.from_numpy(g[f"pooled_{b}"][i]), f"outlier utt{b} hs{i}")
            worst = (min(worst[0], cos), max(worst[1], rel))
    res = model.extract(waves, average=True, want_frames=False, want_pooled=True)
    for b in range(len(lens)):
        check_embedding(res.pooled[b], torch.from_numpy(g[f"meanlast4_pooled_{b}"]), f"outlier utt{b} mean-last-4")
    # the outlier channels dominate both measures above; the ordinary channels on their own (looser: their scale is
    # 1 / 100 of the vector's, and they sit behind LayerNorms whose statistics the outliers set)
    from oracle.make_golden import OUTLIER_CHANNELS
    rest = np.setdiff1d(np.arange(cfg.hidden_size), np.asarray(OUTLIER_CHANNELS))
    worst_rest = (1.0, 0.0)
    for b in range(len(lens)):
        for i in range(cfg.num_hidden_layers + 1):
            a, r = pooled[i, b].cpu()[rest], torch.from_numpy(g[f"pooled_{b}"][i])[rest]
            cos = float(torch.nn.functional.cosine_similarity(a, r, dim=0))
            rel = float((a - r).abs().max() / r.abs().max())
            worst_rest = (min(worst_rest[0], cos), max(worst_rest[1], rel))
    print(f"outlier fixture: worst cosine {worst[0]:.6f}, worst max-rel {worst[1]:.3e}; ordinary channels only: "
          f"cosine {worst_rest[0]:.6f}, max-rel {worst_rest[1]:.3e}")
    assert worst_rest[0] >= 0.995 and worst_rest[1] <= 5e-2, worst_rest
    model.engine.close()


# ------------------------------------------------------------------------------------------------
# batching invariance AT the benched sizes
# ------------------------------------------------------------------------------------------------
def test_benched_wavlm_batch_142_rows_equal_batch_1():
    """bench.py's headline step (142 x 4 s, 28 258 rows: every transformer GEMM on the CTA-pair kernel with the fp32
    residual epilogue): rows 0 / 71 / 141 of the batch equal, bit for bit, the result of encoding that utterance alone
    (where out-proj / FC2 fall to the single-CTA kernel), eager and through the CUDA-graph cache."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    waves = [synth_wave(7000 + j, 64000) for j in range(142)]
    flat = torch.from_numpy(np.concatenate(waves)).cuda()
    big = model.extract_device(flat, [64000] * 142, average=True, want_frames=False, want_pooled=True).pooled.cpu()
    assert big.shape == (142, 1024) and torch.isfinite(big).all()
    for j in (0, 71, 141):
        one = torch.from_numpy(waves[j]).cuda()
        eager = model.extract_device(one, [64000], average=True, use_graph=False).pooled.cpu()[0]
        graphed = model.extract_device(one, [64000], average=True, use_graph=True).pooled.cpu()[0]
        assert torch.equal(eager, big[j]) and torch.equal(graphed, big[j]), j
    # and the golden's 4 s utterance inside the benched batch still matches HF
    g = np.load(os.path.join(REPO, "tests", "golden", "microsoft__wavlm-large.npz"))
    j4 = [int(n) for n in g["lengths"]].index(64000)
    waves[5] = synth_wave(int(g["wave_seed_base"]) + j4, 64000)
    flat = torch.from_numpy(np.concatenate(waves)).cuda()
    big2 = model.extract_device(flat, [64000] * 142, average=True).pooled
    check_embedding(big2[5], torch.from_numpy(g[f"meanlast4_pooled_{j4}"]), "golden utterance inside the 142-batch")


def test_back_to_back_calls_are_bit_identical():
    """The layer-loop kernels are launched with programmatic stream serialization (each sets up under the previous
    kernel's tail and waits before it touches data): twelve encodes of one ragged batch queued without any host
    synchronisation, interleaved with a differently-shaped batch that reuses the same workspace, give the same bits
    every time - for the eager path and for the graph replay of a small batch."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    lens = [64000, 31999, 48000, 16000, 64000, 400, 40001, 64000] * 6
    waves = [synth_wave(9100 + j, n) for j, n in enumerate(lens)]
    flat = torch.from_numpy(np.concatenate(waves)).cuda()
    other = torch.from_numpy(np.concatenate([synth_wave(9300 + j, 24000) for j in range(20)])).cuda()
    outs = []
    for it in range(12):
        outs.append(model.extract_device(flat, lens, average=True, want_frames=False, want_pooled=True, use_graph=False).pooled.clone())
        if it % 3 == 1:
            model.extract_device(other, [24000] * 20, average=True, want_frames=False, want_pooled=True, use_graph=False)
    torch.cuda.synchronize()
    for it in range(1, 12):
        assert torch.equal(outs[it], outs[0]), it
    small = torch.from_numpy(np.concatenate(waves[:8])).cuda()
    ref = model.extract_device(small, lens[:8], average=True, use_graph=False).pooled.clone()
    reps = [model.extract_device(small, lens[:8], average=True, use_graph=True).pooled.clone() for _ in range(12)]
    torch.cuda.synchronize()
    for r in reps:
        assert torch.equal(r, ref)


def test_benched_whisper_batch_32_rows_equal_batch_1():
    """BASELINE configs[1]: Whisper-large-v3, 32 x 30 s. Rows 0 / 17 / 31 of the batch == the window encoded alone."""
    name = "openai/whisper-large-v3"
    cfg, w, model = get_model(name)
    lens = [480000] * 32
    lens[17] = 123457           # a short utterance inside the batch: zero padding, floor-filled mel frames, n_keep crop
    waves = [synth_wave(8000 + j, n) for j, n in enumerate(lens)]
    big = model.extract(waves, average=True, want_frames=False, want_pooled=True).pooled.cpu()
    assert big.shape == (32, 1280) and torch.isfinite(big).all()
    for j in (0, 17, 31):
        one = model.extract([waves[j]], average=True, want_frames=False, want_pooled=True).pooled.cpu()[0]
        assert torch.equal(one, big[j]), j
    _MODELS.pop(name, None)


def test_benched_xlsr_batch_64x8s_rows_equal_batch_1():
    """BASELINE configs[3]: wav2vec2-xls-r-2b, 64 x 8 s per GPU. Rows 0 / 63 == the utterance encoded alone."""
    name = "facebook/wav2vec2-xls-r-2b"
    cfg, w, model = get_model(name)
    waves = [synth_wave(9000 + j, 128000) for j in range(64)]
    big = model.extract(waves, average=True, want_frames=False, want_pooled=True).pooled.cpu()
    assert big.shape == (64, 1920) and torch.isfinite(big).all()
    for j in (0, 63):
        one = model.extract([waves[j]], average=True, want_frames=False, want_pooled=True, use_graph=False).pooled.cpu()[0]
        assert torch.equal(one, big[j]), j
    _MODELS.pop(name, None)


# ------------------------------------------------------------------------------------------------
# boundary additions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny/wavlm", "tiny/whisper"])
def test_int16_pcm_upload_is_bit_identical(name):
    """SERENC_WAV_I16: the kernels scale int16 PCM by 1 / 32768 on load, exactly what librosa hands the reference."""
    cfg, w, model = get_model(name)
    rng = np.random.default_rng(3)
    pcm = [rng.integers(-20000, 20000, size=n).astype(np.int16) for n in (16000, 4001, 48000)]
    flt = [p.astype(np.float32) / np.float32(32768.0) for p in pcm]
    a = model.extract(pcm, average=True, want_frames=True, want_pooled=True)
    b = model.extract(flt, average=True, want_frames=True, want_pooled=True)
    assert torch.equal(a.pooled, b.pooled) and torch.equal(a.packed, b.packed)
    a1 = model.extract(pcm, layer=1, want_frames=False, want_pooled=True, use_graph=False)
    b1 = model.extract(flt, layer=1, want_frames=False, want_pooled=True, use_graph=False)
    assert torch.equal(a1.pooled, b1.pooled)
    mixed = model.extract([pcm[0], flt[1], pcm[2]], average=True, want_frames=True, want_pooled=True)      # mixed kinds: decoded on the host
    assert torch.equal(mixed.pooled, b.pooled)


def test_extract_features_matches_hf_semantics():
    """model(...).extract_features = the feature projection's LayerNorm output (HF modeling_wavlm.py:93-105)."""
    import torch.nn.functional as F
    from interspeech_ser_b200.modeling import AutoFeatureExtractor

    cfg, w, model = get_model("tiny/wavlm")
    proc = AutoFeatureExtractor.from_pretrained("tiny/wavlm", model=model)
    ys = [synth_wave(31, 17777), synth_wave(32, 4001)]
    batch = proc(ys, sampling_rate=16000, return_tensors="pt", padding=True)
    out = model(batch["input_values"], attention_mask=batch["attention_mask"])
    assert out.extract_features.shape == (2, 55, 512) and out["extract_features"] is out.extract_features
    for b, y in enumerate(ys):
        feats = O.conv_feature_encoder(cfg, w, torch.from_numpy(O.normalize_waveform(y)))
        ref = F.layer_norm(feats, (512,), torch.from_numpy(w["featproj.ln.weight"]), torch.from_numpy(w["featproj.ln.bias"]), cfg.layer_norm_eps)
        t = ref.shape[0]
        got = out.extract_features[b, :t].cpu()
        assert float((got - ref).abs().max() / ref.abs().max()) <= 4e-2
        check_embedding(got.mean(0), ref.mean(0), f"extract_features utt{b}")
        assert torch.count_nonzero(out.extract_features[b, t:]) == 0
    one = model(batch["input_values"][1:2], attention_mask=batch["attention_mask"][1:2])     # B == 1 with padding
    assert one.last_hidden_state.shape == (1, 55, cfg.hidden_size)
    assert torch.equal(one.last_hidden_state[0, :12], out.last_hidden_state[1, :12])
    cfg_h, w_h, hub = get_model("tiny/hubert80")
    assert hub(torch.from_numpy(O.normalize_waveform(ys[0]))[None].cuda()).extract_features is None      # HubertModel returns none


@pytest.mark.parametrize("name", ["tiny/wavlm", "tiny/whisper"])
def test_weighted_layer_sum(name):
    """SERENC_REDUCE_WEIGHTED: sum_i softmax(w)_i * hidden_states[i] over the transformer layers (lora_wavlm/model.py:164-181),
    as frames and pooled straight from the layers (the pooled-only path never builds the [sum_T, d] sum)."""
    cfg, w, model = get_model(name)
    L = cfg.num_hidden_layers
    lw = torch.softmax(torch.tensor([0.3, -1.2, 0.7][:L] + [0.1] * max(0, L - 3)), 0).tolist()
    waves = [synth_wave(41, 16000), synth_wave(42, 6000)]
    both = model.extract(waves, layer_weights=lw, want_frames=True, want_pooled=True)
    pooled_only = model.extract(waves, layer_weights=lw, want_frames=False, want_pooled=True)
    for b, y in enumerate(waves):
        if cfg.family == "whisper":
            hs = O.whisper_hidden_states(cfg, w, O.whisper_log_mel(w, y))
            keep = O.whisper_keep_frames(len(y), cfg.hidden_size)
        else:
            hs = O.w2v_hidden_states(cfg, w, y)
            keep = hs[0].shape[0]
        ref = sum(wt * h for wt, h in zip(lw, hs[1:]))[:keep]
        assert both.frames[b].shape == ref.shape
        assert float((both.frames[b].cpu() - ref).abs().max() / ref.abs().max()) <= 4e-2
        check_embedding(both.pooled[b], ref.mean(0), f"{name} weighted utt{b}")
        check_embedding(pooled_only.pooled[b], ref.mean(0), f"{name} weighted pooled-only utt{b}")
        assert torch.allclose(pooled_only.pooled[b], both.pooled[b], atol=2e-5, rtol=1e-4)


def test_pooled_only_mean_of_last4_equals_frame_mean():
    """Mean over layers and mean over frames commute: the pooled-only path (one read of each selected state) against
    the masked mean of the materialised mean-of-last-4 frames."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    waves = [synth_wave(51 + j, n) for j, n in enumerate((64000, 33333, 160000))]
    a = model.extract(waves, average=True, want_frames=True, want_pooled=True)
    b = model.extract(waves, average=True, want_frames=False, want_pooled=True, use_graph=False)
    for j in range(3):
        assert torch.allclose(a.frames[j].mean(0), a.pooled[j], atol=1e-5, rtol=1e-5)
        assert torch.allclose(a.pooled[j], b.pooled[j], atol=2e-5, rtol=1e-4)


def test_request_batching_serves_the_four_thread_call_pattern():
    """INTEGRATION level 1: the reference's ThreadPoolExecutor(4) calls model(**inputs, output_hidden_states=True) with one
    utterance per call (preprocess_speech.py:48-54,120-122). With request batching on, concurrent calls share one packed
    encode and every caller still gets the batch-1 result, bit for bit."""
    from concurrent.futures import ThreadPoolExecutor
    from interspeech_ser_b200.modeling import AutoFeatureExtractor

    cfg, w, model = get_model("tiny/wavlm")
    proc = AutoFeatureExtractor.from_pretrained("tiny/wavlm", model=model)
    ys = [synth_wave(60 + j, 4000 + 977 * j) for j in range(16)]
    inputs = [proc(y, sampling_rate=16000, return_tensors="pt", padding=True) for y in ys]
    ref = [model(**{k: v.to("cuda") for k, v in inp.items()}, output_hidden_states=True) for inp in inputs]
    model.enable_request_batching(max_batch=8, max_wait_ms=20.0)
    try:
        def work(inp):
            out = model(**{k: v.to("cuda") for k, v in inp.items()}, output_hidden_states=True)
            return torch.mean(torch.stack(out.hidden_states[-4:]), dim=0).squeeze(0).cpu(), out
        with ThreadPoolExecutor(max_workers=4) as ex:
            got = list(ex.map(work, inputs))
        q = model._queue
        assert q.requests == 16 and q.batches < 16            # calls were coalesced
    finally:
        model.disable_request_batching()
    for (feats, out), r in zip(got, ref):
        assert len(out.hidden_states) == cfg.num_hidden_layers + 1
        for a, b in zip(out.hidden_states, r.hidden_states):
            assert a.shape == b.shape and torch.equal(a, b)
        assert torch.equal(out.extract_features, r.extract_features)
        assert torch.equal(feats, torch.mean(torch.stack(r.hidden_states[-4:]), dim=0).squeeze(0).cpu())
    # an error in one batch reaches its callers and the dispatcher survives
    model.enable_request_batching(max_batch=4, max_wait_ms=1.0)
    try:
        with pytest.raises(Exception):
            model(torch.zeros(1, 100, device="cuda"))           # shorter than the receptive field
        ok = model(**{k: v.to("cuda") for k, v in inputs[0].items()}, output_hidden_states=True)
        assert torch.equal(ok.last_hidden_state, ref[0].last_hidden_state)
    finally:
        model.disable_request_batching()


def test_graph_cache_replays_equal_eager_calls():
    """extract_device caches a CUDA graph per length signature for small pooled-only calls (BASELINE configs[0]: 8 x 4 s).
    Replays on new waveforms, a second signature, and a return to the first one all equal the eager call."""
    cfg, w, model = get_model("microsoft/wavlm-large")
    model.engine._graphs.clear()
    n0 = 0
    for rep in range(3):
        for lens in ([64000] * 8, [32000, 48000, 16000]):
            wav = torch.from_numpy(np.concatenate([synth_wave(1000 * rep + j, n) for j, n in enumerate(lens)])).cuda()
            a = model.extract_device(wav, lens, average=True)                       # automatic: graph
            b = model.extract_device(wav, lens, average=True, use_graph=False)
            assert torch.equal(a.pooled, b.pooled)
    assert len(model.engine._graphs) == n0 + 2
    big = torch.from_numpy(np.concatenate([synth_wave(j, 64000) for j in range(32)])).cuda()
    model.extract_device(big, [64000] * 32, average=True)                           # 6 368 frames: not launch-bound, stays eager
    assert len(model.engine._graphs) == n0 + 2


def test_poisoned_handle_after_a_kernel_fault():
    """include/serenc.h: kernel faults are sticky. In a child process (the fault kills its CUDA context): a bogus device
    pointer makes a kernel fault, serenc_sync reports SERENC_ERR_CUDA, the handle is poisoned and says so on every later
    call."""
    code = r'''
import sys, ctypes as C, torch
sys.path.insert(0, %r)
from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.engine import Engine
from interspeech_ser_b200.weights import random_init
cfg = configs.get_config("tiny/wavlm")
eng = Engine(cfg, random_init(cfg, 0), 0)
lib = _lib.load_library()
st = torch.cuda.current_stream().cuda_stream
g = torch.ones(128, device="cuda"); out = torch.empty(64, 128, device="cuda")
assert lib.serenc_is_poisoned(eng._h) == 0
rc = lib.serenc_op_layernorm(eng._h, C.c_void_p(0x10), 64, 128, g.data_ptr(), g.data_ptr(), C.c_float(1e-5), 0, out.data_ptr(), None, st)
assert rc == 0, rc                               # the launch itself succeeds: the fault is asynchronous
rc = lib.serenc_sync(eng._h, st)
assert rc == -2, rc                              # SERENC_ERR_CUDA
assert lib.serenc_is_poisoned(eng._h) == 1
rc = lib.serenc_op_layernorm(eng._h, out.data_ptr(), 64, 128, g.data_ptr(), g.data_ptr(), C.c_float(1e-5), 0, out.data_ptr(), None, st)
assert rc == -2 and b"poisoned" in lib.serenc_last_error(), (rc, lib.serenc_last_error())
try:
    eng.synchronize()
except _lib.SerencError as e:
    assert "poisoned" in str(e)
    print("POISONED-OK")
''' % REPO
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "POISONED-OK" in res.stdout, res.stdout + res.stderr


def test_second_replica_on_another_device():
    """the positional-conv kernel's shared-memory opt-in is per device; a second Engine on cuda:1 must work."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from interspeech_ser_b200.modeling import SpeechEncoderModel
    cfg, w, model = get_model("tiny/wavlm")
    other = SpeechEncoderModel(cfg, w, 1)
    waves = [synth_wave(70, 16000), synth_wave(71, 9000)]
    a = model.extract(waves, average=True, want_frames=False).pooled.cpu()
    b = other.extract(waves, average=True, want_frames=False).pooled.cpu()
    assert torch.equal(a, b)
    other.engine.close()


def test_whisper_lora_checkpoint_merged_at_load(tmp_path):
    """preprocess_whisper_pretrained.py:115-138,180-181: the WhisperAudioClassifier state dict (peft r=8, alpha=16 on q_proj /
    v_proj under `whisper.base_model.model.*`) loaded from its .pt gives the embeddings of base + adapter."""
    from interspeech_ser_b200.modeling import AutoModel
    from oracle.make_golden import hf_model

    cfg = configs.get_config("tiny/whisper")
    w = random_init(cfg, 2)
    rng = np.random.default_rng(9)
    sd, merged = {}, {k: v.copy() for k, v in w.items()}
    for k, v in hf_model(cfg, w).state_dict().items():
        mod, _, leaf = k.rpartition(".")
        sd["whisper.base_model.model.encoder." + (f"{mod}.base_layer.{leaf}" if mod.endswith(("q_proj", "v_proj")) else k)] = v
    d = cfg.hidden_size
    for i in range(cfg.num_hidden_layers):
        for s_, name in (("q", "q_proj"), ("v", "v_proj")):
            a = (rng.standard_normal((8, d)) * 0.2).astype(np.float32)
            b = (rng.standard_normal((d, 8)) * 0.2).astype(np.float32)
            mod = f"whisper.base_model.model.encoder.layers.{i}.self_attn.{name}"
            sd[mod + ".lora_A.default.weight"], sd[mod + ".lora_B.default.weight"] = torch.from_numpy(a), torch.from_numpy(b)
            merged[f"layer{i}.{s_}.weight"] = (w[f"layer{i}.{s_}.weight"].astype(np.float64) + 2.0 * (b.astype(np.float64) @ a.astype(np.float64))).astype(np.float32)
    sd["classifier.0.weight"] = torch.zeros(512, d)
    path = str(tmp_path / "whisper_lora_ser.pt")
    torch.save(sd, path)
    model = AutoModel.from_pretrained(path, device=0, config_name="tiny/whisper")
    y = synth_wave(90, 50000)
    res = model.extract([y], layer=-1, want_frames=False, want_pooled=True)
    keep = O.whisper_keep_frames(len(y), cfg.hidden_size)
    mel = O.whisper_log_mel(merged, y)
    check_embedding(res.pooled[0], O.whisper_hidden_states(cfg, merged, mel)[-1][:keep].mean(0), "whisper lora")
    assert float((O.whisper_hidden_states(cfg, w, mel)[-1] - O.whisper_hidden_states(cfg, merged, mel)[-1]).abs().max()) > 1e-2
    model.engine.close()


# ------------------------------------------------------------------------------------------------
# text branch: RoBERTa (preprocessing/preprocess_roberta.py)
# ------------------------------------------------------------------------------------------------
def get_text_model(name):
    if name not in _MODELS:
        from interspeech_ser_b200.text import RobertaModel
        cfg = configs.get_config(name)
        w = random_init(cfg, 0)
        _MODELS[name] = (cfg, w, RobertaModel(cfg, w, 0))
    return _MODELS[name]


def test_roberta_every_hidden_state_vs_oracle_and_hf_golden(golden_dir):
    """Right-padded rows of 80 positions with 80 / 2 / 3 / 17 / 64 / 65 real tokens (a full row, the shortest the
    tokenizer can emit, one and two key blocks): every hidden state, ALL 80 rows of every sequence (the reference saves the
    pad positions too) and the non-pad mean, against the oracle and HF's RobertaModel."""
    from oracle.make_golden import synth_token_rows

    cfg, w, model = get_text_model("tiny/roberta")
    g = load_golden(golden_dir, "tiny/roberta")
    lengths = [int(n) for n in g["lengths"]]
    T = int(g["max_len"])
    rows = synth_token_rows(cfg, int(g["ids_seed"]), lengths, T)
    ids = torch.tensor(rows, dtype=torch.long)
    mask = ids.ne(cfg.pad_token_id).long()
    L = cfg.num_hidden_layers
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True)
    assert len(out.hidden_states) == L + 1 and out.hidden_states[0].shape == (len(rows), T, cfg.hidden_size)
    assert out.last_hidden_state is out["hidden_states"][-1] and out.pooler_output is None
    for b, n in enumerate(lengths):
        hs = O.roberta_hidden_states(cfg, w, rows[b])
        for i in range(L + 1):
            got = out.hidden_states[i][b].cpu()
            assert float((got - hs[i]).abs().max() / hs[i].abs().max()) <= 4e-2, (b, i)       # every row, pads included
            check_embedding(got[:n].mean(0), torch.from_numpy(g[f"pooled_{b}"][i]), f"roberta row{b} hs{i} (tokens) vs HF")
            check_embedding(got.mean(0), torch.from_numpy(g[f"pooled_all_{b}"][i]), f"roberta row{b} hs{i} (all rows) vs HF")
    res = model.extract_tokens(ids, mask, average=True, want_frames=True, want_pooled=True)
    for b, n in enumerate(lengths):
        ref = torch.from_numpy(g[f"meanlast4_{b}"])
        got = res.frames[b][[0, n - 1, T - 1]].cpu()
        assert float((got - ref).abs().max() / ref.abs().max()) <= 4e-2
        assert torch.allclose(res.frames[b][:n].mean(0), res.pooled[b], atol=1e-5, rtol=1e-5)
    # batching invariance: a row alone == the row inside the batch, bit for bit
    alone = model(input_ids=ids[3:4].cuda(), attention_mask=mask[3:4].cuda()).last_hidden_state
    assert torch.equal(alone[0], out.last_hidden_state[3])
    with pytest.raises(IndexError):
        model(input_ids=torch.full((1, 8), cfg.vocab_size, dtype=torch.long).cuda())
    with pytest.raises(NotImplementedError):
        model(input_ids=ids[1:2].cuda(), attention_mask=torch.ones_like(mask[1:2]).cuda())     # mask says "attend to the pads"


def test_roberta_large_vs_hf_golden(golden_dir):
    """Full-size roberta-large (24 post-LN layers, d = 1024, 16 heads of 64; vocabulary 50 265), three rows of 80 positions."""
    from oracle.make_golden import synth_token_rows

    g = load_golden(golden_dir, "roberta-large")
    cfg, w, model = get_text_model("roberta-large")
    lengths = [int(n) for n in g["lengths"]]
    T = int(g["max_len"])
    rows = synth_token_rows(cfg, int(g["ids_seed"]), lengths, T)
    ids = torch.tensor(rows, dtype=torch.long)
    mask = ids.ne(cfg.pad_token_id).long()
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True)
    worst = (1.0, 0.0)
    for b, n in enumerate(lengths):
        for i in range(cfg.num_hidden_layers + 1):
            got = out.hidden_states[i][b]
            cos, rel = check_embedding(got[:n].mean(0), torch.from_numpy(g[f"pooled_{b}"][i]), f"roberta-large row{b} hs{i}")
            check_embedding(got.mean(0), torch.from_numpy(g[f"pooled_all_{b}"][i]), f"roberta-large row{b} hs{i} all rows")
            worst = (min(worst[0], cos), max(worst[1], rel))
        ref = torch.from_numpy(g[f"last_{b}"])
        got = out.last_hidden_state[b][[0, 1, n - 1, T - 1]].cpu()
        assert float((got - ref).abs().max() / ref.abs().max()) <= 4e-2
    print(f"roberta-large worst cosine {worst[0]:.6f}, worst max-rel {worst[1]:.3e}")
    # a realistic batch: 256 rows of 80 through the CTA-pair GEMMs == the same rows in batches of 3
    big_ids = ids.repeat(86, 1)[:256]
    big = model.extract_tokens(big_ids, big_ids.ne(cfg.pad_token_id).long(), layer=-1, want_frames=True).packed.view(256, T, -1)
    for b in range(3):
        assert torch.equal(big[b], out.last_hidden_state[b]) and torch.equal(big[252 + b], out.last_hidden_state[b])
    _MODELS.pop("roberta-large", None)


def test_cli_roberta_contract(tmp_path, capsys):
    """preprocess_roberta.py: CSV (transcription, FileName) -> <basename>.pt of shape [max_len, D] fp32."""
    import json
    import pandas as pd
    from interspeech_ser_b200.cli import main_roberta
    from interspeech_ser_b200.text import RobertaTokenizer
    from test_text_host import CORPUS, train_tiny_bpe

    tok_dir = tmp_path / "tok"
    tok_dir.mkdir()
    vocab, merges = train_tiny_bpe(CORPUS * 3, n_merges=20)
    assert len(vocab) <= configs.get_config("tiny/roberta").vocab_size
    (tok_dir / "vocab.json").write_text(json.dumps(vocab, ensure_ascii=False), encoding="utf-8")
    (tok_dir / "merges.txt").write_text("#version: 0.2\n" + "\n".join(merges) + "\n", encoding="utf-8")
    texts = CORPUS[:6]
    names = [f"MSP-PODCAST_{i:04d}.wav" for i in range(len(texts))]
    csv = tmp_path / "labels.csv"
    pd.DataFrame({"FileName": names, "transcription": texts}).to_csv(csv, index=False)
    out_dir = tmp_path / "feat"
    rc = main_roberta(["--roberta_type", "tiny/roberta", "--df_path", str(csv), "--save_path", str(out_dir), "--random_init",
                       "--tokenizer_path", str(tok_dir), "--max_len", "24", "--use_average", "y", "--batch_texts", "4"])
    assert rc == 0
    cfg, w, _ = get_text_model("tiny/roberta")
    tok = RobertaTokenizer.from_pretrained(str(tok_dir))
    for name, text in zip(names, texts):
        t = torch.load(out_dir / (os.path.splitext(name)[0] + ".pt"))
        assert t.shape == (24, cfg.hidden_size) and t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
        row = tok("nan" if text == "" else text, padding="max_length", truncation=True, max_length=24)["input_ids"]     # pandas reads "" back as NaN
        hs = O.roberta_hidden_states(cfg, w, row)
        ref = torch.stack(hs[-4:]).mean(0)
        assert float((t - ref).abs().max() / ref.abs().max()) <= 4e-2
    capsys.readouterr()
    assert main_roberta(["--roberta_type", "no/such-model", "--df_path", str(csv), "--save_path", str(out_dir), "--tokenizer_path", str(tok_dir)]) == 1
    assert "No pretrained model found with the name no/such-model" in capsys.readouterr().out
    assert main_roberta(["--roberta_type", "tiny/roberta", "--df_path", str(tmp_path / "missing.csv"), "--save_path", str(out_dir)]) == 1
