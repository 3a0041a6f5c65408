"""Dev tool: per-tile clock stamps of the CTA-pair GEMM (MMA issue vs epilogue) for one shape."""
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
M, N = 28416, int(os.environ.get("TRACE_N", "3072"))
CASES = ((256, "none"), (1024, "none"), (1024, "bf16"), (1024, "gelu"), (1024, "resid"))
if os.environ.get("TRACE_RESID_ONLY") == "1":
    CASES = ((1024, "bf16"), (1024, "resid"))
for K, mode in CASES:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out32 = torch.zeros(M, N, device=dev)
    trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    def run():
        if mode == "none":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, None, None, 0, None, None, st))
        elif mode == "bf16":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 0, None, out16.data_ptr(), st))
        elif mode == "gelu":
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, 1, None, out16.data_ptr(), st))
        else:
            _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), out32.data_ptr(), 0, out32.data_ptr(), None, st))
    run(); torch.cuda.synchronize()
    lib.serenc_debug_gemm_trace(eng._h, trace.data_ptr())
    run(); torch.cuda.synchronize()
    lib.serenc_debug_gemm_trace(eng._h, None)
    t = trace.cpu().view(64, 8)
    t0 = int(t[0, 0])
    print(f"--- K={K} {mode}: columns = mma_top, tempty_ok, first_full_ok, mma_committed | epi_top, epi_acc_ready, epi_done (cycles since start)")
    for i in range(10):
        r = [int(v) - t0 if int(v) else -1 for v in t[i, :7]]
        print(f"tile {i:2d}: " + " ".join(f"{v:8d}" for v in r))
