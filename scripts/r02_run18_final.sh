#!/bin/bash
# final production build (TMA epilogue + programmatic dependent launch): full GPU suite, smoke, the driver's bench line, launch list
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02r_pytest.log 2>&1; echo "full suite rc=$?"
tail -3 $O/r02r_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02r_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02r_bench.json 2> $O/r02r_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02r_bench_reference.json 2> $O/r02r_bench_reference.err; echo "reference arm rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --workloads none --no-cpu-baseline"
$CMD > $O/r02r_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 579 -c 400 --csv --log-file $O/r02r_launches_wavlm.csv $CMD > $O/r02r_ncu1.log 2>&1
tail -2 $O/r02r_ncu1.log
