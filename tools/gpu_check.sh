#!/usr/bin/env bash
# Runs the GPU test files as separate processes (a device trap poisons one process only) and
# gathers logs under gpurun_out/.  Usage: gpurun -- 'bash tools/gpu_check.sh [extra cmds]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.csv 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
: > gpurun_out/summary.txt
for f in test_gpu_kernels test_gpu_conv test_gpu_decode; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
