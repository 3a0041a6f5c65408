"""Dev tool: per-CTA clock stamps of the tcgen05 attention kernels (default WavLM-large shape: 142 x 16 heads x T=199;
MODEL=facebook/hubert-xlarge-ls960-ft B=64 T=399 for the wide-head kernel)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.engine import Engine
from interspeech_ser_b200.weights import random_init
dev = torch.device("cuda:0")
cfg = configs.get_config(os.environ.get("MODEL", "microsoft/wavlm-large"))
import dataclasses
cfg = dataclasses.replace(cfg, num_hidden_layers=1)
eng = Engine(cfg, random_init(cfg, 0), 0)
lib = _lib.load_library()
st = torch.cuda.current_stream(dev).cuda_stream
B, T = int(os.environ.get("B", 142)), int(os.environ.get("T", 199))
d = cfg.hidden_size
R = B * T
offs = [i * T for i in range(B + 1)]
qkv = (torch.randn(R, 3 * d, device=dev)).to(torch.bfloat16)
hln = (torch.randn(R, d, device=dev)).to(torch.bfloat16)
out = torch.empty(R, d, device=dev, dtype=torch.bfloat16)
scratch = torch.empty(1 << 16, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wavlm in ((1, 0) if cfg.family == "wavlm" else (0,)):
    def run():
        _lib.check(lib.serenc_op_attention(eng._h, qkv.data_ptr(), _lib.i64_array(offs), B, wavlm, 0, hln.data_ptr() if wavlm else None, out.data_ptr(), scratch.data_ptr(), st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"=== wavlm={wavlm} B={B} T={T}: {min(ts)*1000:.1f} us (min of 5; includes the tiny offsets H2D)")
    if os.environ.get("NOTRACE"):
        continue
    trace = torch.zeros(64 * 48, dtype=torch.int64, device=dev)
    lib.serenc_debug_gemm_trace(eng._h, trace.data_ptr())
    run(); torch.cuda.synchronize()
    lib.serenc_debug_gemm_trace(eng._h, None)
    t = trace.cpu().view(64, 2, 24)
    print("softmax thread0: start setup gate | per block: S_seen p1_done O_prev/rescale_done p2_done+arrive | O_final end dealloc   (cycles since CTA start)")
    print("control thread : start setup Q_in | per block: S_done P_seen V_in(S_next issued) O_done")
    for i in list(range(0, 6)) + list(range(32, 40)):
        for role in (0, 1):
            r = t[i, role]
            t0 = int(r[0])
            if t0 == 0:
                continue
            vals = [int(v) - t0 if int(v) else -1 for v in r[:23]]
            print(f"cta {i:2d} {'smx' if role == 0 else 'ctl'}: " + " ".join(f"{v:6d}" for v in vals[:3]) + " | " +
                  " | ".join(" ".join(f"{v:6d}" for v in vals[4 + 4 * j: 8 + 4 * j]) for j in range(4)) + " || " + " ".join(f"{v:6d}" for v in vals[20:23]))
