set -x
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
# 1. the driver's own command at N = 2: every workload, parity checks, one JSON line
(timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_2gpu.json) 2> gpurun_out/r02g_bench_2gpu.err
tail -c 300 gpurun_out/r02g_bench_2gpu.err; wc -c gpurun_out/r02g_bench_2gpu.json
# 2. strong scaling: ONE fixed corpus, N = 1 then N = 2 on the same box
(timeout 600 python bench.py --workload wavlm-large-corpus --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_corpus_1gpu.json) 2> gpurun_out/r02g_corpus_1gpu.err
(timeout 600 $TR --master-port 29512 bench.py --gpus 2 --workload wavlm-large-corpus --steps 3 --warmup 3 > gpurun_out/r02g_corpus_2gpu.json) 2> gpurun_out/r02g_corpus_2gpu.err
tail -c 300 gpurun_out/r02g_corpus_2gpu.err; wc -c gpurun_out/r02g_corpus_*.json
# 3. the frame-writing CLI on a WAV corpus, one and two ranks
(SERENC_CLI_TIMING=1 timeout 600 python tools/bench_cli.py 2048 2>&1 | grep -E "RESULT|host time") > gpurun_out/r02g_cli_1gpu.log 2>&1
(SERENC_CLI_TIMING=1 timeout 600 $TR --master-port 29513 tools/bench_cli.py 2048 2>&1 | grep -E "RESULT|host time") > gpurun_out/r02g_cli_2gpu.log 2>&1
cat gpurun_out/r02g_cli_1gpu.log gpurun_out/r02g_cli_2gpu.log
