set -x
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --workload wavlm-large-corpus --steps 3 --warmup 3 > gpurun_out/r02h_corpus_8gpu.json) 2> gpurun_out/r02h_corpus_8gpu.err
tail -c 300 gpurun_out/r02h_corpus_8gpu.err
(timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --workload wavlm-large-corpus --steps 3 --warmup 3 > gpurun_out/r02h_corpus_4gpu.json) 2> gpurun_out/r02h_corpus_4gpu.err
(SERENC_CLI_TIMING=1 timeout 600 $TR --nproc-per-node 8 --master-port 29523 tools/bench_cli.py 4096 2>&1 | grep -E "RESULT|host time") > gpurun_out/r02h_cli_8gpu.log 2>&1
cat gpurun_out/r02h_cli_8gpu.log; wc -c gpurun_out/r02h_corpus_*.json
