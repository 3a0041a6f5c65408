#!/bin/bash
# the driver's round-end SCALE command at N = 8 (all workloads of the default line)
set -x
nvidia-smi -L | wc -l
S=$SECONDS
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02m_bench_8gpu.json 2> gpurun_out/r02m_bench_8gpu.err
echo "rc=$? wall_s=$((SECONDS-S))"
wc -c gpurun_out/r02m_bench_8gpu.json
