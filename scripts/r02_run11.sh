set -x
(SERENC_AB_ARMS=1 python -m interspeech_ser_b200.build) > /dev/null 2>&1
export SERENC_AB_ARMS=1
for r in 0 1 2 4 8; do SERENC_LN_RPW=$r timeout 300 python tools/bench_ln.py 2>&1 | grep rows=; done > gpurun_out/r02k_ln_rpw.log 2>&1
cat gpurun_out/r02k_ln_rpw.log
unset SERENC_AB_ARMS
python -m interspeech_ser_b200.build
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -100) > gpurun_out/r02k_pytest.log 2>&1
tail -4 gpurun_out/r02k_pytest.log
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json) 2> gpurun_out/r02k_bench.err
tail -c 400 gpurun_out/r02k_bench.err; wc -c gpurun_out/r02k_bench.json
