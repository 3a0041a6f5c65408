"""Dev tool: which taps of the slab-reuse positional conv come out right (single-tap weights)."""
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
rows, G, cg, taps = 500, 2, 64, 16
g = torch.Generator().manual_seed(1)
x = (torch.randn(rows, G * cg, generator=g) * 0.5).to(torch.bfloat16)
xd = x.to(dev)
res = []
for t0 in range(taps):
    w4 = torch.zeros(G * cg, cg, taps)
    w4[:, :, t0] = torch.randn(G * cg, cg, generator=g) * 0.05
    wp = w4.permute(0, 2, 1).reshape(G * cg, taps * cg).to(torch.bfloat16).to(dev)
    ref = torch.nn.functional.conv1d(x.float().t()[None], w4.to(torch.bfloat16).float(), groups=G)[0].t()
    M = rows - taps + 1
    out = torch.full((M, G * cg), float("nan"), device=dev)
    _lib.check(lib.serenc_op_gemm_grouped(eng._h, xd.data_ptr(), rows, G, cg, taps, wp.data_ptr(), cg, None, 0, out.data_ptr(), st))
    torch.cuda.synchronize()
    o = out.cpu()
    e = float((o - ref).abs().max() / ref.abs().max())
    # does the output match a DIFFERENT shift?
    best = None
    if e > 1e-3:
        for sh in range(-8, 9):
            w5 = torch.zeros_like(w4)
            if 0 <= t0 + sh < taps:
                w5[:, :, t0 + sh] = w4[:, :, t0]
                r2 = torch.nn.functional.conv1d(x.float().t()[None], w5.to(torch.bfloat16).float(), groups=G)[0].t()
                e2 = float((o - r2).abs().max() / r2.abs().max())
                if e2 < 1e-3:
                    best = sh
    res.append((t0, round(e, 4), best))
print("mode", os.environ.get("SERENC_PC_DESC_MODE", "0"), res)
