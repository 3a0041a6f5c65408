#!/bin/bash
# final build on two GPUs: the new back-to-back test, the two-GPU replica test, the driver's scaling command at N = 2
set -x
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "back_to_back or second_replica" > $O/r02s_pytest_n2.log 2>&1; echo "tests rc=$?"
tail -2 $O/r02s_pytest_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02s_bench_2gpu.json 2> $O/r02s_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02s_bench_2gpu.json") if l.startswith("{")][0]
print(d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["parity_ok"], {k: round(v["value"]) for k,v in d["workloads"].items()})
PY
