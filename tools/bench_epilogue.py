"""Dev tool: does the GELU epilogue slow the FC1 GEMM through instruction issue or through the power cap?
Sustained loops of the FC1 shape with epilogue = bf16 cast / exact GELU, SM clock and power sampled by NVML.
(A ReLU control variant was measured once: +1 % time per extra instruction per element under the 1000 W cap, GELU +20 %:
the slowdown is energy. The branch itself cost more than that in code generation and was removed again.)"""
import os, sys, threading, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import pynvml
from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.engine import Engine
from interspeech_ser_b200.weights import random_init
dev = torch.device("cuda:0")
cfg = configs.get_config("tiny/wavlm")
eng = Engine(cfg, random_init(cfg, 0), 0)
lib = _lib.load_library()
st = torch.cuda.current_stream(dev).cuda_stream
pynvml.nvmlInit()
hd = pynvml.nvmlDeviceGetHandleByIndex(0)
M, N, K = 28416, 4096, 1024
a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
def run(act):
    if act < 0:
        torch.matmul(a, w.t(), out=out16)
    else:
        _lib.check(lib.serenc_op_gemm(eng._h, a.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr(), None, act, None, out16.data_ptr(), st))
for name, act in (("bf16", 0), ("gelu", 1), ("cublas", -1), ("bf16", 0), ("gelu", 1)):
    for _ in range(5):
        run(act)
    torch.cuda.synchronize()
    samples, stop = [], False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1000.0))
            time.sleep(0.02)
    th = threading.Thread(target=sampler); th.start()
    n = 8000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run(act)
    e1.record(); torch.cuda.synchronize()
    stop = True; th.join()
    t = e0.elapsed_time(e1) / n
    samples = samples[len(samples) // 2:]   # second half: clocks and the power reading have settled
    clk = sorted(s[0] for s in samples)[len(samples) // 2]; pw = sorted(s[1] for s in samples)[len(samples) // 2]
    print(f"{name:6s}: {t*1000:7.1f} us  {2.0*M*N*K/(t*1e-3)/1e12:7.1f} TFLOP/s   SM {clk} MHz  {pw:.0f} W  ({len(samples)} samples)  -> {2.0*M*N*K/(t*1e-3)/1e12/clk*1000:.1f} TFLOP/s per GHz", flush=True)
