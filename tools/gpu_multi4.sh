#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 300 --warmup 5 > $O/r02w_bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"; tail -n 3 $O/bench_n4.err
python tools/summarize_bench.py $O/r02w_bench_n4.json
