"""Dev tool: time the tcgen05 GEMM entry point against torch.matmul (cuBLAS) on a list of shapes.
    python tools/bench_gemm.py            (SERENC_FORCE_1CTA=1 to force the single-CTA kernel)"""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.engine import Engine
from interspeech_ser_b200.weights import random_init

dev = torch.device("cuda:0")
cfg = configs.get_config("tiny/wavlm")
eng = Engine(cfg, random_init(cfg, 0), 0)
lib = _lib.load_library()
st = torch.cuda.current_stream(dev).cuda_stream
shapes = [(8192, 8192, 8192), (25472, 3072, 1024), (25472, 1024, 1024), (25472, 4096, 1024), (25472, 1024, 4096), (25472, 3072, 8192),
          (48000, 3840, 1280), (48000, 5120, 1280), (48000, 1280, 5120), (1592, 3072, 1024), (1592, 4096, 1024)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K) in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out32 = torch.zeros(M, N, device=dev)
    res = {}
    def run(mode):
        if mode == "bf16":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 0, None, out16.data_ptr(), st))
        elif mode == "gelu":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 1, None, out16.data_ptr(), st))
        elif mode == "resid":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), out32.data_ptr(), 0, out32.data_ptr(), None, st))
        else:
            torch.matmul(a, w.t(), out=out16)
    for mode in ("bf16", "gelu", "resid", "cublas"):
        for _ in range(3):
            run(mode)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(mode); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        res[mode] = 2.0 * M * N * K / (t * 1e-3) / 1e12
    print(f"M={M:6d} N={N:5d} K={K:5d}  " + "  ".join(f"{k}={v:7.1f}" for k, v in res.items()), flush=True)
