set -x
nvidia-smi -L | wc -l; free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# the driver's own command at N = 8: every workload, parity checks, one JSON line (host memory / time check)
(/usr/bin/time -v timeout 900 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02l_bench_8gpu.json) 2> gpurun_out/r02l_bench_8gpu.err
grep -E "Elapsed|Maximum resident" gpurun_out/r02l_bench_8gpu.err; free -g | head -2
# strong scaling with the final build: 8, 4, 2, 1 ranks on the same box
for n in 8 4 2 1; do
  (timeout 600 $TR --nproc-per-node $n --master-port $((29540 + n)) bench.py --gpus $n --workload wavlm-large-corpus --steps 3 --warmup 3 > gpurun_out/r02l_corpus_${n}gpu.json) 2> gpurun_out/r02l_corpus_${n}gpu.err
done
wc -c gpurun_out/r02l_*.json
