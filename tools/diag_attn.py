"""Dev tool: attention error per utterance length / head / query tile (tcgen05 path, no bias)."""
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
d, H = cfg.hidden_size, cfg.num_attention_heads
def ref(qkv, offs):
    out = torch.zeros(qkv.shape[0], d)
    for b in range(len(offs) - 1):
        s, e = offs[b], offs[b + 1]
        q, k, v = [qkv[s:e, i * d:(i + 1) * d].float().view(e - s, H, 64).transpose(0, 1) for i in range(3)]
        a = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1) @ v
        out[s:e] = a.transpose(0, 1).reshape(e - s, d)
    return out
for lens in ([16], [32], [33], [64], [96], [128], [129], [160], [199], [333], [199, 1, 64, 65, 333, 12], [199] * 40):
    offs = [0]
    for t in lens: offs.append(offs[-1] + t)
    R = offs[-1]
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(R, 3 * d, generator=g).to(torch.bfloat16)
    out = torch.full((R, d), float("nan"), dtype=torch.bfloat16, device=dev)
    scratch = torch.empty(1 << 16, dtype=torch.uint8, device=dev)
    qd = qkv.to(dev)
    _lib.check(lib.serenc_op_attention(eng._h, qd.data_ptr(), _lib.i64_array(offs), len(lens), 0, 0, None, out.data_ptr(), scratch.data_ptr(), st))
    torch.cuda.synchronize()
    r = ref(qkv, offs)
    o = out.float().cpu()
    errs = []
    for b in range(len(lens)):
        for h in range(H):
            for t0 in range(0, lens[b], 128):
                a = o[offs[b] + t0: min(offs[b] + t0 + 128, offs[b + 1]), h * 64:(h + 1) * 64]
                bb = r[offs[b] + t0: min(offs[b] + t0 + 128, offs[b + 1]), h * 64:(h + 1) * 64]
                e = float((a - bb).abs().max() / bb.abs().max()) if torch.isfinite(a).all() else float("nan")
                if not (e < 2e-2):
                    errs.append((b, h, t0, round(e, 3)))
    print(f"lens={lens if len(lens) < 8 else str(lens[:2]) + '...x' + str(len(lens))}: bad tiles {errs[:12]}{' ...' if len(errs) > 12 else ''} ({len(errs)} bad)")
