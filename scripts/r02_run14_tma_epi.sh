#!/bin/bash
# fp32 GEMM epilogue through TMA: parity, same-box A/B (development build carries the switch), full suite, bench
set -x
export SERENC_AB_ARMS=1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm" > $O/r02n_pytest_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -3 $O/r02n_pytest_gemm.log
timeout 300 python tools/bench_gemm.py > $O/r02n_gemm_tma.log 2>&1; echo rc=$?
SERENC_GEMM_NO_TMA_EPI=1 timeout 300 python tools/bench_gemm.py > $O/r02n_gemm_reg.log 2>&1; echo rc=$?
paste -d'\n' $O/r02n_gemm_tma.log $O/r02n_gemm_reg.log | grep "M="
TRACE_N=1024 TRACE_RESID_ONLY=1 timeout 120 python tools/trace_gemm.py > $O/r02n_trace_tma.log 2>&1
TRACE_N=1024 TRACE_RESID_ONLY=1 SERENC_GEMM_NO_TMA_EPI=1 timeout 120 python tools/trace_gemm.py > $O/r02n_trace_reg.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02n_pytest.log 2>&1; echo "full suite rc=$?"
tail -3 $O/r02n_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --workloads none --no-cpu-baseline > $O/r02n_bench_tma_$i.json 2> $O/r02n_bench_tma_$i.err; echo rc=$?
SERENC_GEMM_NO_TMA_EPI=1 timeout 300 python bench.py --steps 20 --warmup 5 --workloads none --no-cpu-baseline > $O/r02n_bench_reg_$i.json 2> $O/r02n_bench_reg_$i.err; echo rc=$?
done
python - <<'PY'
import json
for n in ("tma_1","reg_1","tma_2","reg_2"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02n_bench_{n}.json") if l.startswith("{")][0]
        kb=d["kernel_breakdown"]
        print(n, round(d["value"]), round(d["ms_per_step"],3), "out", round(kb["gemm_out"]["ms_per_step"],3), "fc2", round(kb["gemm_fc2"]["ms_per_step"],3), d["parity_ok"])
    except Exception as e: print(n, "failed", e)
PY
