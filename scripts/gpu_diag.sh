#!/bin/bash
# Bring-up diagnostics on a B200 box: every stage in its own process under a timeout; logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/diag_gpu.txt 2>&1
for t in "$@"; do
  name=$(echo "$t" | tr ' /' '__')
  echo "=== $t ===" | tee -a gpurun_out/diag_all.log
  timeout 600 python tools/diag_gpu.py $t > gpurun_out/diag_$name.log 2>&1
  echo "exit=$?" >> gpurun_out/diag_$name.log
  tail -n 60 gpurun_out/diag_$name.log | tee -a gpurun_out/diag_all.log
done
