"""Dev tool: isolate mainloop vs epilogue cost of the CTA-pair GEMM (K sweep, with and without output)."""
import os, sys
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M, N = 28416, 3072
for K in (256, 512, 1024, 2048, 4096):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out32 = torch.zeros(M, N, device=dev)
    res = {}
    def run(mode):
        if mode == "none":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, None, None, 0, None, None, st))
        elif mode == "bf16":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 0, None, out16.data_ptr(), st))
        elif mode == "gelu":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 1, None, out16.data_ptr(), st))
        elif mode == "f32":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 0, out32.data_ptr(), None, st))
        elif mode == "resid":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), out32.data_ptr(), 0, out32.data_ptr(), None, st))
        else:
            torch.matmul(a, w.t(), out=out16)
    for mode in ("none", "bf16", "gelu", "f32", "resid", "cublas"):
        for _ in range(3):
            run(mode)
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(mode); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        res[mode] = (2.0 * M * N * K / (t * 1e-3) / 1e12, t * 1e3)
    print(f"M={M} N={N} K={K:5d}  " + "  ".join(f"{k}={v[0]:7.1f}TF/{v[1]:6.0f}us" for k, v in res.items()), flush=True)
