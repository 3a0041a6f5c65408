set -x
nvidia-smi -L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -80) > gpurun_out/r02a_pytest.log 2>&1
(timeout 60 tools/mufu_cost.bin) > gpurun_out/r02a_mufu.log 2>&1
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json) 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_mufu.log; tail -c 1500 gpurun_out/r02a_bench.err; wc -c gpurun_out/r02a_bench.json
