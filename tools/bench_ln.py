"""Dev tool: LayerNorm pass timings on the model's shapes (op entry, L2 flushed, median of 7).
    python tools/bench_ln.py        (development build: SERENC_LN_RPW=1|2|4|8 selects the rows-per-warp variant)"""
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
for rows, cols, gelu in ((28258, 1024, 0), (48000, 1280, 0), (25536, 1920, 0), (908000, 512, 1), (454000, 512, 1)):
    x = torch.randn(rows, cols, device=dev)
    g = torch.ones(cols, device=dev); b = torch.zeros(cols, device=dev)
    o16 = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
    def run():
        _lib.check(lib.serenc_op_layernorm(eng._h, x.data_ptr(), rows, cols, g.data_ptr(), b.data_ptr(), 1e-5, gelu, None, o16.data_ptr(), st))
    for _ in range(3):
        run()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[3]
    print(f"rows={rows:7d} cols={cols:5d} gelu={gelu}: {t*1000:8.1f} us  {rows*cols*6/(t*1e-3)/1e9:7.0f} GB/s (fp32 in, bf16 out)  RPW={os.environ.get('SERENC_LN_RPW','default')}", flush=True)
