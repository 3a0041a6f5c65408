#!/bin/bash
# TMA epilogue: ring depth 3 (4 operand stages) against depth 2 (5 stages), and the register path for K > 2048
set -x
export SERENC_AB_ARMS=1
O=gpurun_out
SERENC_GEMM_TMA_NB=3 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm" > $O/r02o_pytest_gemm_nb3.log 2>&1; echo "gemm tests nb3 rc=$?"
tail -2 $O/r02o_pytest_gemm_nb3.log
SERENC_GEMM_TMA_NB=3 timeout 300 python tools/bench_gemm.py > $O/r02o_gemm_nb3.log 2>&1; echo rc=$?
SERENC_GEMM_TMA_NB=2 timeout 300 python tools/bench_gemm.py > $O/r02o_gemm_nb2.log 2>&1; echo rc=$?
paste -d'\n' $O/r02o_gemm_nb3.log $O/r02o_gemm_nb2.log | grep "M="
TRACE_N=1024 TRACE_RESID_ONLY=1 SERENC_GEMM_TMA_NB=3 timeout 120 python tools/trace_gemm.py > $O/r02o_trace_nb3.log 2>&1
grep -A7 resid $O/r02o_trace_nb3.log
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --workloads none --no-cpu-baseline > $O/r02o_bench_$n.json 2> $O/r02o_bench_$n.err; echo "$n rc=$?"
}
for i in 1 2; do
run nb3_$i SERENC_GEMM_TMA_NB=3
run nb3k_$i SERENC_GEMM_TMA_NB=3 SERENC_GEMM_TMA_MAX_KB=32
run nb2k_$i SERENC_GEMM_TMA_NB=2 SERENC_GEMM_TMA_MAX_KB=32
run nb2_$i SERENC_GEMM_TMA_NB=2
run reg_$i SERENC_GEMM_NO_TMA_EPI=1
done
python - <<'PY'
import json
for n in ("nb3","nb3k","nb2k","nb2","reg"):
  for i in (1,2):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02o_bench_{n}_{i}.json") if l.startswith("{")][0]
        kb=d["kernel_breakdown"]
        print(n, i, round(d["value"]), round(d["ms_per_step"],3), "out", round(kb["gemm_out"]["ms_per_step"],3), "fc2", round(kb["gemm_fc2"]["ms_per_step"],3), d["parity_ok"], d["clocks"]["sm_mhz"])
    except Exception as e: print(n, "failed", e)
PY
