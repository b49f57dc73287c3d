#!/bin/bash
# Staged GPU validation: safe kernels first, tcgen05 in its own process (a faulting kernel poisons its CUDA context).
# Usage (under gpurun): bash tools/gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --maxfail=4 --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "== $name exit $? =="; tail -n 25 "gpurun_out/$name.log"
}
run safe 600 tests/test_gpu_ops.py -m gpu -k "simt or decode or mask or empty or launch"
run tc_first 300 tests/test_gpu_ops.py -m gpu -k "tcgen05 and 1x128x128x2x1x128"
run tc 900 tests/test_gpu_ops.py -m gpu -k "tcgen05 or agree"
run modules 600 tests/test_gpu_modules.py -m gpu
run full 900 tests/test_gpu_full_size.py -m gpu
