"""Host-side logic (CPU): configs, weight conversion, scheduler, audio ingest, HF-shaped output objects."""
import os

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import audio_io, configs, scheduler
from interspeech_ser_b200.engine import layer_mask_of
from interspeech_ser_b200.modeling import ModelOutput
from interspeech_ser_b200.weights import fold_weight_norm, from_hf_state_dict, random_init, slaney_mel_filters, whisper_sinusoids


def test_model_constants():
    c = configs.get_config("microsoft/wavlm-large")
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size, c.head_dim) == (1024, 24, 16, 4096, 64)
    assert not c.conv_bias and c.family == "wavlm"
    h = configs.get_config("facebook/hubert-xlarge-ls960-ft")
    assert (h.hidden_size, h.num_hidden_layers, h.head_dim, h.intermediate_size) == (1280, 48, 80, 5120)
    x = configs.get_config("wav2vec2-xls-r-2b")
    assert (x.hidden_size, x.num_hidden_layers, x.head_dim, x.intermediate_size) == (1920, 48, 120, 7680)
    w = configs.get_config("openai/whisper-large-v3")
    assert (w.hidden_size, w.num_hidden_layers, w.num_attention_heads, w.num_mel_bins) == (1280, 32, 20, 128)
    with pytest.raises(OSError):  # the reference catches OSError from from_pretrained (preprocess_speech.py:115-117)
        configs.get_config("not/a-model")


def test_layer_mask_semantics():
    assert layer_mask_of([-1], 24) == (1 << 24, [24])
    assert layer_mask_of([0, 3], 24) == (0b1001, [0, 3])
    assert layer_mask_of([-4, -3, -2, -1], 24)[1] == [21, 22, 23, 24]
    with pytest.raises(IndexError):
        layer_mask_of([25], 24)   # hidden_states[25] on 25 states: the reference's swallowed IndexError (defect D1)


def test_fold_weight_norm_matches_torch():
    conv = torch.nn.Conv1d(32, 32, 16, padding=8, groups=4)
    conv = torch.nn.utils.parametrizations.weight_norm(conv, name="weight", dim=2)
    with torch.no_grad():
        conv.parametrizations.weight.original0.mul_(1.7)
    w = fold_weight_norm(conv.parametrizations.weight.original0, conv.parametrizations.weight.original1)
    np.testing.assert_allclose(w, conv.weight.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_random_init_shapes_and_determinism():
    cfg = configs.get_config("tiny/wavlm")
    a, b = random_init(cfg, 0), random_init(cfg, 0)
    assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
    assert a["conv0.weight"].shape == (512, 1, 10) and a["conv3.weight"].shape == (512, 512, 3)
    assert a["posconv.weight"].shape == (128, 32, 16) and a["rel_attn_embed"].shape == (320, 2)
    assert a["layer1.gru.weight"].shape == (8, 64) and "conv0.bias" not in a
    w = random_init(configs.get_config("tiny/whisper"), 0)
    assert "layer0.k.bias" not in w and w["embed_positions"].shape == (1500, 128) and w["mel_filters"].shape == (201, 80)


def test_hf_state_dict_conversion_roundtrip():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    pytest.importorskip("transformers")
    from oracle.make_golden import hf_model
    for name in ("tiny/wavlm", "tiny/wav2vec2", "tiny/whisper", "tiny/wavlm-base", "tiny/hubert-base"):
        cfg = configs.get_config(name)
        w = random_init(cfg, 1)
        back = from_hf_state_dict(cfg, hf_model(cfg, w).state_dict())
        assert sorted(back) == sorted(w), name
        for k in w:
            np.testing.assert_allclose(back[k], w[k], rtol=1e-5, atol=1e-6, err_msg=f"{name}:{k}")


def test_lora_merge_equals_adapter_forward(tmp_path):
    """peft-layout state dict (preprocess_speech_pretrained.py:120-130: r=8, alpha=16, q_proj/v_proj) -> dense weights
    whose Linear output equals base(x) + (alpha/r) * B(A(x)); classifier head dropped; file loading path."""
    from interspeech_ser_b200.weights import load_checkpoint_dir, merge_lora
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path_root)
    pytest.importorskip("transformers")
    from oracle.make_golden import hf_model
    cfg = configs.get_config("tiny/wavlm")
    w = random_init(cfg, 2)
    sd = hf_model(cfg, w).state_dict()
    g = torch.Generator().manual_seed(0)
    peft_sd, expect = {}, {}
    for k, v in sd.items():
        mod = k.rsplit(".", 1)[0]
        if mod.endswith(("q_proj", "v_proj")):
            peft_sd[f"wavlm.base_model.model.{mod}.base_layer.{k.rsplit('.', 1)[1]}"] = v
            if k.endswith(".weight"):
                a = torch.randn(8, v.shape[1], generator=g) * 0.1
                b = torch.randn(v.shape[0], 8, generator=g) * 0.1
                peft_sd[f"wavlm.base_model.model.{mod}.lora_A.default.weight"] = a
                peft_sd[f"wavlm.base_model.model.{mod}.lora_B.default.weight"] = b
                expect[k] = (a, b, v)
        else:
            peft_sd[f"wavlm.base_model.model.{k}"] = v
    peft_sd["classifier.0.weight"] = torch.zeros(4, 4)
    merged = merge_lora(peft_sd)
    assert sorted(merged) == sorted(sd)
    x = torch.randn(5, cfg.hidden_size, generator=g)
    for k, (a, b, v) in expect.items():
        want = x @ v.T + 2.0 * ((x @ a.T) @ b.T)
        got = x @ torch.from_numpy(np.asarray(merged[k])).T
        assert torch.allclose(got, want, atol=1e-5, rtol=1e-5), k
    path = str(tmp_path / "lora_ser.pt")
    torch.save(peft_sd, path)
    canon = load_checkpoint_dir(cfg, path)
    assert sorted(canon) == sorted(w)
    assert not np.allclose(canon["layer0.q.weight"], w["layer0.q.weight"]) and np.array_equal(canon["layer0.k.weight"], w["layer0.k.weight"])
    bad = dict(peft_sd)
    del bad["wavlm.base_model.model.encoder.layers.0.attention.q_proj.lora_B.default.weight"]
    with pytest.raises(ValueError):
        merge_lora(bad)


def test_mel_filters_and_sinusoids():
    fb = slaney_mel_filters(128)
    assert fb.shape == (201, 128) and fb.min() >= 0 and (fb.max(axis=0) > 0).all()
    nnz = (fb != 0).sum(axis=0)
    assert nnz.min() >= 1 and nnz.max() <= 12          # sparse triangles: the kernel stores them as CSR
    pos = whisper_sinusoids(1500, 1280)
    assert pos.shape == (1500, 1280) and np.allclose(pos[0, :640], 0) and np.allclose(pos[0, 640:], 1)


def test_wav_bytes_decode_matches_file_decode(tmp_path):
    """The CLI reads files in I/O threads (read_bytes) and decodes in one thread (load_audio_bytes): same samples as
    load_audio on the path, for 16-bit PCM, a float32 WAVE and an odd trailing byte; 16-bit scaling is exact."""
    x = (np.random.default_rng(2).standard_normal(16001) * 0.3).astype(np.float32)
    p = str(tmp_path / "a.wav")
    audio_io.write_wav(p, x)
    y_file, sr = audio_io.load_audio(p)
    y_bytes, sr2 = audio_io.load_audio_bytes(audio_io.read_bytes(p), p)
    assert sr == sr2 == 16000 and np.array_equal(y_file, y_bytes) and y_bytes.dtype == np.float32
    pcm = np.clip(np.round(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    assert np.array_equal(y_bytes, pcm.astype(np.float32) / 32768.0)
    data = audio_io.read_bytes(p)
    y_odd, _ = audio_io.decode_wav(data[:-1], "truncated")          # data chunk one byte short: whole samples only
    assert len(y_odd) == len(x) - 1 and np.array_equal(y_odd, y_bytes[:-1])
    import struct
    f32 = x.astype("<f4").tobytes()
    wav = (b"RIFF" + struct.pack("<I", 36 + len(f32)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32) +
           b"data" + struct.pack("<I", len(f32)) + f32)
    yf, _ = audio_io.decode_wav(wav, "float")
    assert np.array_equal(yf, x)
    with pytest.raises(ValueError):
        audio_io.decode_wav(b"garbage", "broken")


def test_shard_by_cost_partitions_files_and_balances():
    """File-level rank sharding of the CLI (cost = size on disk): a partition, deterministic, LPT-balanced, and each
    rank's list ascending in cost so that its decode windows hold similar lengths."""
    rng = np.random.default_rng(3)
    costs = [float(v) for v in rng.integers(64_000, 640_000, size=1000)]
    for world in (1, 2, 8):
        assign = scheduler.shard_by_cost(costs, world)
        assert sorted(i for a in assign for i in a) == list(range(1000))
        assert scheduler.shard_by_cost(costs, world) == assign
        loads = [sum(costs[i] for i in a) for a in assign]
        assert max(loads) - min(loads) <= max(costs)
        for a in assign:
            assert [costs[i] for i in a] == sorted(costs[i] for i in a)
    assert scheduler.shard_by_cost([], 4) == [[], [], [], []]
    assert scheduler.shard_by_cost([5.0], 2) == [[0], []]


def test_scheduler_batches_cover_every_utterance_once():
    cfg = configs.get_config("microsoft/wavlm-large")
    rng = np.random.default_rng(0)
    lens = [int(v) for v in rng.integers(32000, 320000, size=500)]
    batches = scheduler.make_batches(cfg, lens, frame_budget=16384)
    seen = sorted(i for b in batches for i in b.indices)
    assert seen == list(range(500))
    assert all(b.frames <= 16384 for b in batches)
    for b in batches:  # length-sorted: bucketed by construction
        ls = [lens[i] for i in b.indices]
        assert ls == sorted(ls)
    assign = scheduler.shard_batches(batches, 8)
    assert sorted(j for a in assign for j in a) == list(range(len(batches)))
    loads = [sum(batches[j].flops for j in a) for a in assign]
    assert max(loads) <= 1.25 * (sum(loads) / 8) + max(b.flops for b in batches)
    assert scheduler.shard_batches(batches, 8) == assign  # deterministic
    with pytest.raises(ValueError):
        scheduler.make_batches(cfg, [64000, 399])


def test_flop_model_matches_survey():
    # WavLM-large 4 s = 147.3 GFLOP, Whisper-large-v3 window = 2273.8 GFLOP
    assert abs(scheduler.utterance_flops(configs.get_config("wavlm-large"), 64000) / 1e9 - 147.3) < 0.1
    assert abs(scheduler.utterance_flops(configs.get_config("whisper-large-v3"), 1) / 1e9 - 2273.8) < 0.1


def test_merge_rank_results():
    merged = scheduler.merge_rank_results([{0: "a", 2: "c"}, {1: "b"}], 3)
    assert merged == ["a", "b", "c"]
    with pytest.raises(ValueError):
        scheduler.merge_rank_results([{0: "a"}, {0: "b"}], 1)
    with pytest.raises(ValueError):
        scheduler.merge_rank_results([{0: "a"}], 2)


def test_wav_roundtrip_and_resample(tmp_path):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(16000) * 0.1).astype(np.float32)
    p = str(tmp_path / "a.wav")
    audio_io.write_wav(p, x, 16000)
    y, sr = audio_io.load_audio(p)
    assert sr == 16000 and y.dtype == np.float32 and np.abs(y - x).max() <= 1.0 / 32768 + 1e-7
    p8 = str(tmp_path / "b.wav")
    t = np.arange(8000) / 8000.0
    audio_io.write_wav(p8, 0.5 * np.sin(2 * np.pi * 440 * t), 8000)
    z, sr = audio_io.load_audio(p8)
    assert sr == 16000 and abs(len(z) - 16000) <= 1
    with pytest.raises(ValueError):
        (tmp_path / "c.wav").write_bytes(b"not a wav file at all")
        audio_io.read_wav(str(tmp_path / "c.wav"))


def test_model_output_access_patterns():
    hs = (torch.zeros(1, 2, 3), torch.ones(1, 2, 3))
    out = ModelOutput(last_hidden_state=hs[-1], extract_features=None, hidden_states=hs)
    assert out.hidden_states is hs and out["hidden_states"] is hs      # both forms appear in preprocess_speech.py:56,67
    assert out.last_hidden_state is hs[-1] and out[0] is hs[-1]
    assert out.extract_features is None and "extract_features" not in out


def test_cli_parsers_keep_reference_flags():
    from interspeech_ser_b200.cli import build_parser
    for whisper in (False, True):
        a = build_parser(whisper).parse_args([])
        assert (a.seed, a.ssl_type, a.save_path, a.wav_dir, a.num_workers, a.n_layer, a.use_average) == (7, "wavlm-large", "./", "./", 4, -1, "n")


def test_epilogue_gelu_constants_are_exact_to_roundoff():
    """The polynomial-exp2 erfc used by the CUDA epilogues (csrc/common.cuh) against the exact erf GELU."""
    import re
    from oracle import gelu_fit
    assert gelu_fit.max_error() < 1e-6
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "interspeech_ser_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_erf_fast(float x)"):]
    body = body[:body.index("return fmaf(-a, e, hx + a);")]
    consts = [float(v) for v in re.findall(r"(-?\d\.\d{9}e[+-]\d\d)f", body)]
    assert consts == list(reversed(gelu_fit.COEFFS))      # Horner order in the kernel = descending powers


def test_request_coalescing_queue_groups_concurrent_calls():
    """enable_request_batching (SURVEY 8b "Threading"): concurrent single-utterance calls are collected by the dispatcher
    and handed to ONE engine call per kind; every caller gets its own slice; an exception reaches exactly the callers of
    the failing group and the dispatcher keeps serving. (Host logic only: a stand-in model records what it is given.)"""
    import threading
    import time
    from concurrent.futures import ThreadPoolExecutor

    from interspeech_ser_b200.modeling import _CoalescingQueue

    class FakeModel:
        def __init__(self):
            self.calls = []

        def _run_coalesced(self, kind, items):
            self.calls.append((kind, len(items)))
            if any(it[1] == "boom" for it in items):
                raise ValueError("bad utterance")
            time.sleep(0.01)
            return [(kind, it[1]) for it in items]

    m = FakeModel()
    q = _CoalescingQueue(m, max_batch=8, max_wait_ms=30.0)
    try:
        gate = threading.Barrier(4)

        def work(i):
            gate.wait()
            return q.submit((i % 2 == 0, f"utt{i}")).result(timeout=10)
        with ThreadPoolExecutor(max_workers=4) as ex:
            got = list(ex.map(work, range(4)))
        assert got == [(True, "utt0"), (False, "utt1"), (True, "utt2"), (False, "utt3")]
        assert q.requests == 4 and q.batches == 2 and sorted(m.calls) == [(False, 2), (True, 2)]   # one call per kind
        f_bad, f_ok = q.submit((True, "boom")), q.submit((False, "fine"))
        with pytest.raises(ValueError):
            f_bad.result(timeout=10)
        assert f_ok.result(timeout=10) == (False, "fine")
        assert q.submit((True, "after")).result(timeout=10) == (True, "after")
    finally:
        q.close()
    with pytest.raises(RuntimeError):
        q.submit((True, "closed"))


def test_int16_wav_decode_is_a_view_and_matches_float_decode(tmp_path):
    """audio_io keep_int16: mono 16-bit PCM comes back as the int16 samples (the GPU applies 1 / 32768, include/serenc.h
    SERENC_WAV_I16); scaled on the host it equals the float decode bit for bit; other formats fall back to float32."""
    from interspeech_ser_b200 import audio_io
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(5000) * 0.2).astype(np.float32)
    path = str(tmp_path / "a.wav")
    audio_io.write_wav(path, x)
    data = audio_io.read_bytes(path)
    pcm, sr = audio_io.decode_wav(data, path, keep_int16=True)
    flt, _ = audio_io.decode_wav(data, path)
    assert pcm.dtype == np.int16 and sr == 16000 and pcm.shape == flt.shape
    assert np.array_equal(pcm.astype(np.float32) * np.float32(1.0 / 32768.0), flt)
    y, sr2 = audio_io.load_audio_bytes(data, path, sr=8000, keep_int16=True)       # resampled: float32 again
    assert y.dtype == np.float32 and sr2 == 8000 and abs(len(y) - 2500) <= 1
