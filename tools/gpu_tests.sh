#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel kills the CUDA context of that process only),
# each under its own timeout. Output -> gpurun_out/tests_*.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --timeout 300 -x > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
done
