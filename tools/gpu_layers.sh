#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 300 > $O/r02w_bench_n1.json 2> $O/bench_w.err; echo "bench rc=$?"; tail -n 3 $O/bench_w.err
python tools/summarize_bench.py $O/r02w_bench_n1.json | tail -8
bash tools/gpu_tcsrc.sh
