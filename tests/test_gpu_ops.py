"""Op-level known-answer tests on the GPU, through the C ABI, against a plain fp32 PyTorch statement of each op
(operands rounded to bf16 first, so the tolerance is accumulate-order only). Run with `-m gpu` on a B200."""
import sys

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.weights import random_init

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def bf(x):
    return x.to(torch.bfloat16).contiguous()


@pytest.fixture(scope="module")
def ctx():
    from interspeech_ser_b200.engine import Engine
    cfg = configs.get_config("tiny/wavlm")
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    return cfg, w, eng, _lib.load_library(), torch.device("cuda:0")


def stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 256, 512), (1592, 3072, 1024), (1592, 1024, 4096), (999, 1920, 1920),
                                   (130, 1280, 1280), (77, 200, 192), (1, 1024, 1024), (25472, 1024, 1024),
                                   (5000, 1920, 512), (9999, 5120, 1280), (20000, 264, 320)])
def test_gemm_epilogues(ctx, M, N, K):
    cfg, w, eng, lib, dev = ctx
    g = torch.Generator().manual_seed(M * 7 + N)
    a = bf(torch.randn(M, K, generator=g) * 0.5).to(dev)
    wt = bf(torch.randn(N, K, generator=g) * 0.05).to(dev)
    bias = (torch.randn(N, generator=g) * 0.1).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev)
    base = a.float() @ wt.float().t()
    # plain fp32
    out = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, None, None, 0, out.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, base) < 1e-4
    # bias + exact GELU -> bf16
    o16 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), None, 1, None, o16.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(o16, torch.nn.functional.gelu(base + bias)) < 6e-3   # bf16 output rounding (2^-8)
    # bias + fp32 residual, in place
    out.copy_(resid)
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), out.data_ptr(), 0, out.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, base + bias + resid) < 1e-4


@pytest.mark.parametrize("M,N,K", [(3000, 1056, 256), (10240, 1024, 128), (2561, 2048, 320)])
def test_gemm_fp32_tma_epilogue_variants(ctx, M, N, K):
    """The CTA-pair GEMM's fp32 epilogue moves residual and output blocks by TMA: row tail (M not a multiple of 256 / 32),
    column tail (N = 1056: the last tile column holds one 32-column block), residual out of place, GELU before the
    residual add, no bias; rows past M and the untouched source stay as they were."""
    cfg, w, eng, lib, dev = ctx
    g = torch.Generator().manual_seed(M + N + K)
    a = bf(torch.randn(M, K, generator=g) * 0.5).to(dev)
    wt = bf(torch.randn(N, K, generator=g) * 0.05).to(dev)
    bias = (torch.randn(N, generator=g) * 0.1).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev)
    base = a.float() @ wt.float().t()
    # residual out of place, one guard row behind the output
    out = torch.full((M + 1, N), 7.0, device=dev)
    keep = resid.clone()
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), resid.data_ptr(), 0, out.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out[:M], base + bias + resid) < 1e-4
    assert torch.equal(resid, keep) and bool((out[M] == 7.0).all())
    # GELU, then the residual, in place, no bias
    out2 = resid.clone()
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, None, out2.data_ptr(), 1, out2.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out2, torch.nn.functional.gelu(base) + resid) < 1e-3   # tanh-form GELU with MUFU.TANH (2^-11 relative)
    # bias only (no residual: nothing is loaded, the block is built in shared memory and stored)
    out3 = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, wt.data_ptr(), N, bias.data_ptr(), None, 0, out3.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out3, base + bias) < 1e-4


@pytest.mark.parametrize("rows_in,C,N,taps,s", [(1000, 512, 512, 3, 2), (1001, 512, 512, 2, 2), (777, 128, 256, 3, 1),
                                                 (3002, 1280, 1280, 3, 2), (4100, 512, 512, 3, 2), (40001, 512, 512, 3, 2),
                                                 (40000, 512, 512, 2, 2), (96064, 128, 1280, 3, 1)])
def test_implicit_gemm_conv(ctx, rows_in, C, N, taps, s):
    """Strided Conv1d over channels-last rows == GEMM over (tap, channel) with row offsets instead of im2col."""
    cfg, w, eng, lib, dev = ctx
    g = torch.Generator().manual_seed(rows_in)
    x = bf(torch.randn(rows_in, C, generator=g) * 0.5).to(dev)
    wt3 = torch.randn(N, C, taps, generator=g) * 0.03
    wt = bf(wt3.permute(0, 2, 1).reshape(N, taps * C)).to(dev)
    M = (rows_in - taps) // s + 1
    ref = torch.nn.functional.conv1d(x.float().t()[None], bf(wt3).float().to(dev), stride=s)[0].t()
    out = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.serenc_op_gemm(eng._h, x.data_ptr(), M, taps * C, s * C, wt.data_ptr(), N, None, None, 0, out.data_ptr(), None, stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-4


@pytest.mark.parametrize("rows,G,cg,cg_pad,taps", [(500, 16, 64, 64, 128), (300, 4, 32, 64, 16), (260, 8, 80, 128, 16), (300, 16, 120, 128, 15)])
def test_grouped_positional_conv(ctx, rows, G, cg, cg_pad, taps):
    cfg, w, eng, lib, dev = ctx
    g = torch.Generator().manual_seed(rows + G)
    xv = torch.randn(rows, G, cg, generator=g) * 0.5
    x = torch.zeros(rows, G, cg_pad)
    x[:, :, :cg] = xv
    x = bf(x.reshape(rows, G * cg_pad)).to(dev)
    w4 = torch.randn(G * cg, cg, taps, generator=g) * 0.05
    wp = torch.zeros(G * cg, taps, cg_pad)
    wp[:, :, :cg] = w4.permute(0, 2, 1)
    wp = bf(wp.reshape(G * cg, taps * cg_pad)).to(dev)
    bias = (torch.randn(G * cg, generator=g) * 0.1).to(dev)
    ref = torch.nn.functional.conv1d(bf(xv).float().reshape(rows, G * cg).t()[None].to(dev), bf(w4).float().to(dev), bias, groups=G)[0].t()
    M = rows - taps + 1
    out = torch.full((M, G * cg), float("nan"), device=dev)
    _lib.check(lib.serenc_op_gemm_grouped(eng._h, x.data_ptr(), rows, G, cg_pad, taps, wp.data_ptr(), cg, bias.data_ptr(), 0, out.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-4


@pytest.mark.parametrize("cols", [128, 256, 512, 640, 1024, 1280, 1920])
def test_layernorm(ctx, cols):
    cfg, w, eng, lib, dev = ctx
    g = torch.Generator().manual_seed(cols)
    rows = 1001
    x = (torch.randn(rows, cols, generator=g) * 2 + 0.3).to(dev)
    ga = (1 + 0.1 * torch.randn(cols, generator=g)).to(dev)
    be = (0.1 * torch.randn(cols, generator=g)).to(dev)
    ref = torch.nn.functional.layer_norm(x, (cols,), ga, be, 1e-5)
    o32 = torch.empty_like(x)
    o16 = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
    _lib.check(lib.serenc_op_layernorm(eng._h, x.data_ptr(), rows, cols, ga.data_ptr(), be.data_ptr(), 1e-5, 0, o32.data_ptr(), o16.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert float((o32 - ref).abs().max()) < 5e-6
    assert float((o16.float() - ref).abs().max()) < 2.5e-2   # bf16 rounding of values up to ~6
    _lib.check(lib.serenc_op_layernorm(eng._h, x.data_ptr(), rows, cols, ga.data_ptr(), be.data_ptr(), 1e-5, 1, None, o16.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert float((o16.float() - torch.nn.functional.gelu(ref)).abs().max()) < 2.5e-2


def _attn_ref(qkv, offs, H, bias_fn=None):
    d = qkv.shape[1] // 3
    dh = d // H
    out = torch.zeros(qkv.shape[0], d)
    for b in range(len(offs) - 1):
        s, e = offs[b], offs[b + 1]
        T = e - s
        q, k, v = [qkv[s:e, i * d:(i + 1) * d].float().view(T, H, dh).transpose(0, 1) for i in range(3)]
        sc = (q @ k.transpose(1, 2)) * dh ** -0.5
        if bias_fn is not None:
            sc = sc + bias_fn(s, e)
        out[s:e] = (torch.softmax(sc, -1) @ v).transpose(0, 1).reshape(T, d)
    return out


def test_attention_wavlm_long_utterance():
    """T = 1100 and 1500 frames (22 s / 30 s): relative positions beyond the 1023-entry Toeplitz table are clamped
    (buckets saturate at |delta| >= 778, so the clamp is exact) and the bias window of a query tile outgrows the
    4-CTA/SM shared-memory budget; next to a short utterance in the same packed batch."""
    from interspeech_ser_b200.engine import Engine
    from oracle import ssl_oracle as O
    lib, dev = _lib.load_library(), torch.device("cuda:0")
    cfg = configs.get_config("tiny/wavlm")
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    d, H = cfg.hidden_size, cfg.num_attention_heads
    g = torch.Generator().manual_seed(11)
    lens = [1100, 37, 1500]
    offs = [0]
    for t in lens:
        offs.append(offs[-1] + t)
    R = offs[-1]
    qkv = bf(torch.randn(R, 3 * d, generator=g))
    hln = bf(torch.randn(R, d, generator=g))
    scratch = torch.empty(4096, dtype=torch.uint8, device=dev)
    out = torch.full((R, d), float("nan"), dtype=torch.bfloat16, device=dev)
    li = 0

    def bias_fn(s, e):
        T = e - s
        pb = O.wavlm_position_bias(cfg, w, T)
        xh = hln[s:e].float().view(T, H, d // H).transpose(0, 1)
        proj = torch.nn.functional.linear(xh, torch.from_numpy(w[f"layer{li}.gru.weight"]), torch.from_numpy(w[f"layer{li}.gru.bias"]))
        gate = torch.sigmoid(proj.view(H, T, 2, 4).sum(-1))
        gg = gate[..., 0] * (gate[..., 1] * torch.from_numpy(w[f"layer{li}.gru.const"]).view(H, 1) - 1.0) + 2.0
        return gg[:, :, None] * pb
    _lib.check(lib.serenc_op_attention(eng._h, qkv.to(dev).data_ptr(), _lib.i64_array(offs), len(lens), 1, li, hln.to(dev).data_ptr(),
                                       out.data_ptr(), scratch.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, _attn_ref(qkv, offs, H, bias_fn)) < 6e-3


@pytest.mark.parametrize("name", ["tiny/wavlm", "tiny/wav2vec2", "tiny/hubert80", "tiny/w2v120"])
def test_attention_varlen(name):
    """Packed variable-length attention (ragged: T = 1, 64, 65, 199, 333, 12), head_dim 64 / 80 / 120, and WavLM's gated
    relative-position bias against the oracle's explicit [H, T, T] bias."""
    from interspeech_ser_b200.engine import Engine
    from oracle import ssl_oracle as O
    lib, dev = _lib.load_library(), torch.device("cuda:0")
    cfg = configs.get_config(name)
    w = random_init(cfg, 0)
    eng = Engine(cfg, w, 0)
    d, H = cfg.hidden_size, cfg.num_attention_heads
    g = torch.Generator().manual_seed(4)
    lens = [199, 1, 64, 65, 333, 12]
    offs = [0]
    for t in lens:
        offs.append(offs[-1] + t)
    R = offs[-1]
    qkv = bf(torch.randn(R, 3 * d, generator=g))
    hln = bf(torch.randn(R, d, generator=g))
    scratch = torch.empty(4096, dtype=torch.uint8, device=dev)
    out = torch.full((R, d), float("nan"), dtype=torch.bfloat16, device=dev)
    qd, hd_ = qkv.to(dev), hln.to(dev)
    _lib.check(lib.serenc_op_attention(eng._h, qd.data_ptr(), _lib.i64_array(offs), len(lens), 0, 0, None, out.data_ptr(), scratch.data_ptr(), stream(dev)))
    torch.cuda.synchronize()
    assert rel_err(out, _attn_ref(qkv, offs, H)) < 6e-3
    if cfg.family == "wavlm":
        li = 1

        def bias_fn(s, e):
            T = e - s
            pb = O.wavlm_position_bias(cfg, w, T)
            xh = hln[s:e].float().view(T, H, d // H).transpose(0, 1)
            proj = torch.nn.functional.linear(xh, torch.from_numpy(w[f"layer{li}.gru.weight"]), torch.from_numpy(w[f"layer{li}.gru.bias"]))
            gate = torch.sigmoid(proj.view(H, T, 2, 4).sum(-1))
            gg = gate[..., 0] * (gate[..., 1] * torch.from_numpy(w[f"layer{li}.gru.const"]).view(H, 1) - 1.0) + 2.0
            return gg[:, :, None] * pb
        out.fill_(float("nan"))
        _lib.check(lib.serenc_op_attention(eng._h, qd.data_ptr(), _lib.i64_array(offs), len(lens), 1, li, hd_.data_ptr(), out.data_ptr(), scratch.data_ptr(), stream(dev)))
        torch.cuda.synchronize()
        assert rel_err(out, _attn_ref(qkv, offs, H, bias_fn)) < 6e-3
