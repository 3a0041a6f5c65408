set -x
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -100) > gpurun_out/r02i_pytest.log 2>&1
tail -6 gpurun_out/r02i_pytest.log
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json) 2> gpurun_out/r02i_bench.err
tail -c 400 gpurun_out/r02i_bench.err; wc -c gpurun_out/r02i_bench.json
