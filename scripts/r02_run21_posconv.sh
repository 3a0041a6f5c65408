#!/bin/bash
# positional conv without the zero K steps / with N rounded to 16 (HuBERT-xlarge: 80-channel groups): suite + the two wide-model workloads
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02v_pytest.log 2>&1; echo "suite rc=$?"
tail -3 $O/r02v_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 --workloads hubert-xlarge --no-cpu-baseline > $O/r02v_bench.json 2> $O/r02v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02v_bench.json") if l.startswith("{")][0]
print("wavlm", round(d["value"]), round(d["ms_per_step"],2), "posconv", round(d["kernel_breakdown"]["gemm_posconv"]["ms_per_step"],3), d["parity_ok"], d["clocks"]["sm_mhz"])
for k,v in d["workloads"].items():
    print(k, round(v["value"]), round(v["ms_per_step"],2), "posconv", round(v["kernel_breakdown"]["gemm_posconv"]["ms_per_step"],3), round(v["kernel_breakdown"]["gemm_posconv"]["tflops"]), v["parity_check"]["ok"])
PY
