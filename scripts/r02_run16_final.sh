#!/bin/bash
# production build with the TMA epilogue: full GPU suite, the driver's bench line, ncu launch list + --set full of one layer
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02p_pytest.log 2>&1; echo "full suite rc=$?"
tail -3 $O/r02p_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02p_bench.json 2> $O/r02p_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02p_smoke.log 2>&1; echo "smoke rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --workloads none --no-cpu-baseline"
$CMD > $O/r02p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 579 -c 400 --csv --log-file $O/r02p_launches_wavlm.csv $CMD > $O/r02p_ncu1.log 2>&1
$CMD > $O/r02p_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05_2cta_kernel|layernorm_rows_kernel" -s 540 -c 16 -o /tmp/r02p_prof $CMD > $O/r02p_ncu2.log 2>&1
ncu -i /tmp/r02p_prof.ncu-rep --page raw --csv > $O/r02p_ncu_wavlm_raw.csv 2> /dev/null
ls -la $O/ | tail -12
tail -2 $O/r02p_ncu2.log
