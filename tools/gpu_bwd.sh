#!/bin/bash
mkdir -p gpurun_out
C=$PWD/vats_multimodal_lm_b200/csrc
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_bwd.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_bwd.log
for lib in libvats_attn libvats_noocc libvats_attn libvats_noocc; do
  echo "== $lib"
  VATS_ATTN_LIB=$C/$lib.so timeout 200 python tools/run_backward.py --time | grep forward
  VATS_ATTN_LIB=$C/$lib.so timeout 200 python tools/run_backward.py 4 4096 32 8 128 -1 --time | grep forward
  VATS_ATTN_LIB=$C/$lib.so timeout 200 python tools/run_backward.py 64 196 16 8 72 -1 --time | grep forward
done
