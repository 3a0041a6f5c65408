#!/bin/bash
# whole GPU suite on a two-GPU box (runs the second-replica test that one-GPU boxes skip)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest_2gpu_box.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r02t_pytest_2gpu_box.log
