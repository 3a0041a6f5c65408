#!/bin/bash
# programmatic dependent launch for the layer-loop kernels: same-box A/B (development build carries SERENC_NO_PDL)
set -x
export SERENC_AB_ARMS=1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_round2.py -m gpu -x -q > $O/r02q_pytest_part.log 2>&1; echo "tests rc=$?"
tail -2 $O/r02q_pytest_part.log
run() { n=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --workloads wavlm-large-c1,hubert-xlarge,whisper-large-v3 --no-cpu-baseline > $O/r02q_bench_$n.json 2> $O/r02q_bench_$n.err; echo "$n rc=$?"
}
for i in 1 2; do
run pdl_$i SERENC_DUMMY=1
run nopdl_$i SERENC_NO_PDL=1
done
python - <<'PY'
import json
for n in ("pdl_1","nopdl_1","pdl_2","nopdl_2"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02q_bench_{n}.json") if l.startswith("{")][0]
        w=d["workloads"]
        print(n, "wavlm", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "| c1", round(w["wavlm-large-c1"]["ms_per_step"],3), "| hubert", round(w["hubert-xlarge"]["ms_per_step"],2), "| whisper", round(w["whisper-large-v3"]["ms_per_step"],2), d["parity_ok"], d["clocks"]["sm_mhz"])
    except Exception as e: print(n, "failed", e)
PY
