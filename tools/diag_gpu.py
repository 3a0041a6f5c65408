"""GPU diagnostics: op-level and stage-level error reports against torch fp32 / the oracle.

Not a pytest module — a bring-up tool that prints numbers instead of asserting, one sub-command per process so
that a device-side trap in one kernel cannot poison the others:

    python tools/diag_gpu.py gemm | conv | grouped | ln | attn | model <cfg> | logmel | whisper <cfg>
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from interspeech_ser_b200 import _lib, configs  # noqa: E402
from interspeech_ser_b200.engine import Engine, REDUCE_NONE, REDUCE_MEAN  # noqa: E402
from interspeech_ser_b200.weights import random_init  # noqa: E402

DEV = torch.device("cuda:0")


def stream():
    return torch.cuda.current_stream(DEV).cuda_stream


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def tiny_engine(name="tiny/wavlm"):
    cfg = configs.get_config(name)
    w = random_init(cfg, 0)
    return cfg, w, Engine(cfg, w, 0)


def bf(x):
    return x.to(torch.bfloat16).contiguous()


def cmd_gemm():
    cfg, w, eng = tiny_engine()
    lib = _lib.load_library()
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [(128, 64, 64), (128, 128, 64), (256, 256, 128), (300, 256, 512), (1592, 3072, 1024), (1592, 1024, 4096),
              (1592, 4096, 1024), (999, 1920, 1920), (25472, 1024, 1024), (130, 1280, 1280), (77, 200, 192), (5000, 512, 1536)]
    for (M, N, K) in shapes:
        for mode in ("plain", "bias+gelu->bf16", "bias+resid->f32"):
            a = bf(torch.randn(M, K, generator=g) * 0.5).to(DEV)
            wt = bf(torch.randn(N, K, generator=g) * 0.05).to(DEV)
            bias = (torch.randn(N, generator=g) * 0.1).to(DEV)
            resid = torch.randn(M, N, generator=g).to(DEV)
            ref = a.float() @ wt.float().t()
            out32 = torch.full((M, N), float("nan"), device=DEV)
            out16 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
            if mode == "plain":
                st = lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, None, None, 0, out32.data_ptr(), None, stream())
                got = out32
            elif mode == "bias+gelu->bf16":
                ref = torch.nn.functional.gelu(ref + bias)
                st = lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), None, 1, None, out16.data_ptr(), stream())
                got = out16
            else:
                ref = ref + bias + resid
                out32.copy_(resid)
                st = lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), out32.data_ptr(), 0, out32.data_ptr(), None, stream())
                got = out32
            _lib.check(st)
            torch.cuda.synchronize()
            e = rel_err(got, ref)
            nan = int(torch.isnan(got.float()).sum())
            print(f"gemm M={M} N={N} K={K} {mode:18s} rel_err={e:.3e} nan={nan} {'OK' if e < 1.5e-2 and nan == 0 else 'FAIL'}", flush=True)


def cmd_conv():
    cfg, w, eng = tiny_engine()
    lib = _lib.load_library()
    g = torch.Generator(device="cpu").manual_seed(1)
    for (rows_in, Cc, N, taps, s) in [(1000, 512, 512, 3, 2), (1001, 512, 512, 2, 2), (777, 128, 256, 3, 1), (3002, 1280, 1280, 3, 2), (4100, 512, 512, 3, 2)]:
        x = bf(torch.randn(rows_in, Cc, generator=g) * 0.5).to(DEV)
        wt3 = torch.randn(N, Cc, taps, generator=g) * 0.03          # torch conv layout [N, C, taps]
        wt = bf(wt3.permute(0, 2, 1).reshape(N, taps * Cc)).to(DEV)  # tap-major K
        M = (rows_in - taps) // s + 1
        ref = torch.nn.functional.conv1d(x.float().t()[None], bf(wt3).float().to(DEV), stride=s)[0].t()
        out = torch.full((M, N), float("nan"), device=DEV)
        st = lib.serenc_op_gemm(eng._h, x.data_ptr(), M, taps * Cc, s * Cc, wt.data_ptr(), N, None, None, 0, out.data_ptr(), None, stream())
        _lib.check(st)
        torch.cuda.synchronize()
        e = rel_err(out, ref)
        print(f"conv rows={rows_in} C={Cc} N={N} taps={taps} s={s} M={M} rel_err={e:.3e} {'OK' if e < 1.5e-2 else 'FAIL'}", flush=True)


def cmd_grouped():
    cfg, w, eng = tiny_engine()
    lib = _lib.load_library()
    g = torch.Generator(device="cpu").manual_seed(2)
    for (rows, G, cg, cg_pad, taps, npg) in [(500, 16, 64, 64, 128, 64), (300, 4, 32, 64, 16, 32), (260, 8, 80, 128, 16, 80), (300, 16, 120, 128, 15, 120)]:
        x = torch.zeros(rows, G * cg_pad)
        xv = torch.randn(rows, G, cg, generator=g) * 0.5
        x.view(rows, G, cg_pad)[:, :, :cg] = xv
        x = bf(x).to(DEV)
        w4 = torch.randn(G * npg, cg, taps, generator=g) * 0.05
        wp = torch.zeros(G * npg, taps, cg_pad)
        wp[:, :, :cg] = w4.permute(0, 2, 1)
        wp = bf(wp.reshape(G * npg, taps * cg_pad)).to(DEV)
        bias = (torch.randn(G * npg, generator=g) * 0.1).to(DEV)
        xin = bf(xv).float().reshape(rows, G * cg).t()[None].to(DEV)
        ref = torch.nn.functional.conv1d(xin, bf(w4).float().to(DEV), bias, groups=G)[0].t()
        M = rows - taps + 1
        out = torch.full((M, G * npg), float("nan"), device=DEV)
        st = lib.serenc_op_gemm_grouped(eng._h, x.data_ptr(), rows, G, cg_pad, taps, wp.data_ptr(), npg, bias.data_ptr(), 0, out.data_ptr(), stream())
        _lib.check(st)
        torch.cuda.synchronize()
        e = rel_err(out, ref)
        print(f"grouped rows={rows} G={G} cg={cg} taps={taps} rel_err={e:.3e} {'OK' if e < 1.5e-2 else 'FAIL'}", flush=True)


def cmd_ln():
    cfg, w, eng = tiny_engine()
    lib = _lib.load_library()
    g = torch.Generator(device="cpu").manual_seed(3)
    for cols in (128, 256, 512, 640, 1024, 1280, 1920):
        rows = 1001
        x = (torch.randn(rows, cols, generator=g) * 2 + 0.3).to(DEV)
        ga = (1 + 0.1 * torch.randn(cols, generator=g)).to(DEV)
        be = (0.1 * torch.randn(cols, generator=g)).to(DEV)
        ref = torch.nn.functional.layer_norm(x, (cols,), ga, be, 1e-5)
        o32 = torch.empty_like(x)
        o16 = torch.empty(rows, cols, device=DEV, dtype=torch.bfloat16)
        _lib.check(lib.serenc_op_layernorm(eng._h, x.data_ptr(), rows, cols, ga.data_ptr(), be.data_ptr(), 1e-5, 0, o32.data_ptr(), o16.data_ptr(), stream()))
        torch.cuda.synchronize()
        print(f"ln cols={cols} f32 err={float((o32 - ref).abs().max()):.3e} bf16 err={float((o16.float() - ref).abs().max()):.3e}", flush=True)
        _lib.check(lib.serenc_op_layernorm(eng._h, x.data_ptr(), rows, cols, ga.data_ptr(), be.data_ptr(), 1e-5, 1, None, o16.data_ptr(), stream()))
        torch.cuda.synchronize()
        print(f"   +gelu bf16 err={float((o16.float() - torch.nn.functional.gelu(ref)).abs().max()):.3e}", flush=True)


def attn_ref(qkv, offs, H, bias_fn=None):
    d = qkv.shape[1] // 3
    dh = d // H
    out = torch.zeros(qkv.shape[0], d)
    for b in range(len(offs) - 1):
        s, e = offs[b], offs[b + 1]
        T = e - s
        q, k, v = [qkv[s:e, i * d:(i + 1) * d].float().view(T, H, dh).transpose(0, 1) for i in range(3)]
        sc = (q @ k.transpose(1, 2)) * dh ** -0.5
        if bias_fn is not None:
            sc = sc + bias_fn(b, s, e)
        out[s:e] = (torch.softmax(sc, -1) @ v).transpose(0, 1).reshape(T, d)
    return out


def cmd_attn():
    lib = _lib.load_library()
    sys.path.insert(0, REPO)
    from oracle import ssl_oracle as O
    for name in ("tiny/wavlm", "tiny/wav2vec2", "tiny/hubert80", "tiny/w2v120"):
        cfg, w, eng = tiny_engine(name)
        d, H = cfg.hidden_size, cfg.num_attention_heads
        g = torch.Generator(device="cpu").manual_seed(4)
        lens = [199, 1, 64, 65, 333, 12]
        offs = [0]
        for t in lens:
            offs.append(offs[-1] + t)
        R = offs[-1]
        qkv = bf(torch.randn(R, 3 * d, generator=g))
        hln = bf(torch.randn(R, d, generator=g))
        scratch = torch.empty(4096, dtype=torch.uint8, device=DEV)
        out = torch.full((R, d), float("nan"), dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.serenc_op_attention(eng._h, qkv.to(DEV).data_ptr(), _lib.i64_array(offs), len(lens), 0, 0, None, out.data_ptr(), scratch.data_ptr(), stream()))
        torch.cuda.synchronize()
        ref = attn_ref(qkv, offs, H)
        print(f"attn[{name}] plain hd={d // H} rel_err={rel_err(out, ref):.3e} nan={int(torch.isnan(out.float()).sum())}", flush=True)
        if cfg.family == "wavlm":
            def bias_fn(b, s, e):
                T = e - s
                pb = O.wavlm_position_bias(cfg, w, T)
                xh = hln[s:e].float().view(T, H, d // H).transpose(0, 1)
                proj = torch.nn.functional.linear(xh, torch.from_numpy(w["layer1.gru.weight"]), torch.from_numpy(w["layer1.gru.bias"]))
                gate = torch.sigmoid(proj.view(H, T, 2, 4).sum(-1))
                gg = gate[..., 0] * (gate[..., 1] * torch.from_numpy(w["layer1.gru.const"]).view(H, 1) - 1.0) + 2.0
                return gg[:, :, None] * pb
            hl = hln.to(DEV)
            out.fill_(float("nan"))
            _lib.check(lib.serenc_op_attention(eng._h, qkv.to(DEV).data_ptr(), _lib.i64_array(offs), len(lens), 1, 1, hl.data_ptr(), out.data_ptr(), scratch.data_ptr(), stream()))
            torch.cuda.synchronize()
            ref = attn_ref(qkv, offs, H, bias_fn)
            print(f"attn[{name}] wavlm-gated rel_err={rel_err(out, ref):.3e} nan={int(torch.isnan(out.float()).sum())}", flush=True)


def synth_wave(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * np.float32(0.0886)).astype(np.float32)


def cmd_model(name, lens=None, full_layers=True):
    from oracle import ssl_oracle as O
    cfg = configs.get_config(name)
    t0 = time.time()
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    print(f"model[{name}] weights+engine {time.time() - t0:.1f}s", flush=True)
    if isinstance(lens, str):
        lens = [int(v) for v in lens.split(",")]
    lens = lens or [400, 401, 719, 720, 4001, 17777, 32000]
    waves = [synth_wave(7 + j, n) for j, n in enumerate(lens)]
    starts, off = [], 0
    for wv in waves:
        starts.append(off)
        off += len(wv)
    wav = torch.from_numpy(np.concatenate(waves)).to(DEV)
    L = cfg.num_hidden_layers
    frames, pooled, offs, idx = eng.encode_w2v(wav, starts, lens, normalize=True, layers=range(L + 1), reduce=REDUCE_NONE, want_frames=True, want_pooled=True)
    torch.cuda.synchronize()
    frames = frames.cpu()
    pooled = pooled.cpu()
    for b, wv in enumerate(waves):
        hs = O.w2v_hidden_states(cfg, w, wv)
        s, e = offs[b], offs[b + 1]
        errs = [rel_err(frames[i, s:e], hs[i]) for i in range(L + 1)]
        cos = [float(torch.nn.functional.cosine_similarity(pooled[i, b], hs[i].mean(0), dim=0)) for i in range(L + 1)]
        show = list(range(L + 1)) if L <= 4 else [0, 1, 2, L // 2, L - 1, L]
        print(f"  utt{b} len={lens[b]} T={e - s}: " + " ".join(f"hs{i}:err={errs[i]:.2e},cos={cos[i]:.5f}" for i in show), flush=True)
    fm, pm, _, _ = eng.encode_w2v(wav, starts, lens, normalize=True, layers=list(range(max(0, L - 3), L + 1)), reduce=REDUCE_MEAN, want_frames=True, want_pooled=True)
    torch.cuda.synchronize()
    ref = frames[max(0, L - 3):].mean(0)
    print(f"  mean-last-4 frames vs own hidden states: {rel_err(fm.cpu(), ref):.2e}", flush=True)


def cmd_logmel():
    from oracle import ssl_oracle as O
    from oracle.make_golden import logmel_signals
    cfg = configs.get_config("tiny/whisper128")
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    sigs = logmel_signals()
    names = list(sigs)
    waves = [sigs[n] for n in names]
    starts, off = [], 0
    for wv in waves:
        starts.append(off)
        off += len(wv)
    wav = torch.from_numpy(np.concatenate(waves)).to(DEV)
    t0 = time.time()
    mel = eng.logmel(wav, starts, [len(x) for x in waves])
    torch.cuda.synchronize()
    mel = mel.cpu()
    for b, n in enumerate(names):
        ref = O.whisper_log_mel(w, waves[b])
        print(f"logmel[{n}] max abs err = {float((mel[b] - ref).abs().max()):.3e}", flush=True)


def cmd_whisper(name):
    from oracle import ssl_oracle as O
    cfg = configs.get_config(name)
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    lens = [16000, 80000, 480000]
    waves = [synth_wave(7 + j, n) for j, n in enumerate(lens)]
    starts, off = [], 0
    for wv in waves:
        starts.append(off)
        off += len(wv)
    wav = torch.from_numpy(np.concatenate(waves)).to(DEV)
    mel = eng.logmel(wav, starts, lens)
    L = cfg.num_hidden_layers
    keep = [O.whisper_keep_frames(n, cfg.hidden_size) for n in lens]
    frames, pooled, idx = eng.encode_whisper(mel, layers=range(L + 1), reduce=REDUCE_NONE, n_keep=keep, want_frames=True, want_pooled=True)
    torch.cuda.synchronize()
    frames, pooled = frames.cpu(), pooled.cpu()
    for b, wv in enumerate(waves):
        m_ref = O.whisper_log_mel(w, wv)
        hs = O.whisper_hidden_states(cfg, w, m_ref)
        errs = [rel_err(frames[i, b * 1500:(b + 1) * 1500], hs[i]) for i in range(L + 1)]
        cos = [float(torch.nn.functional.cosine_similarity(pooled[i, b], hs[i][:keep[b]].mean(0), dim=0)) for i in range(L + 1)]
        show = list(range(L + 1)) if L <= 4 else [0, 1, L // 2, L]
        print(f"  whisper[{name}] utt{b} len={lens[b]}: mel_err={float((mel[b].cpu() - m_ref).abs().max()):.2e} " + " ".join(f"hs{i}:err={errs[i]:.2e},cos={cos[i]:.5f}" for i in show), flush=True)


if __name__ == "__main__":
    cmd = sys.argv[1]
    args = sys.argv[2:]
    {"gemm": cmd_gemm, "conv": cmd_conv, "grouped": cmd_grouped, "ln": cmd_ln, "attn": cmd_attn, "model": cmd_model,
     "logmel": cmd_logmel, "whisper": cmd_whisper}[cmd](*args)
