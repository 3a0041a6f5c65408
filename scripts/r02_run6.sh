set -x
(timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/r02f_ops.log 2>&1
tail -3 gpurun_out/r02f_ops.log
(timeout 600 python tools/bench_gemm.py) > gpurun_out/r02f_gemm_microbench.log 2>&1
cat gpurun_out/r02f_gemm_microbench.log
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -100) > gpurun_out/r02f_pytest.log 2>&1
tail -4 gpurun_out/r02f_pytest.log
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json) 2> gpurun_out/r02f_bench.err
tail -c 400 gpurun_out/r02f_bench.err; wc -c gpurun_out/r02f_bench.json
