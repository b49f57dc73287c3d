#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 300 --warmup 5 > $O/r02z_bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -n 3 $O/bench_n8.err
python tools/summarize_bench.py $O/r02z_bench_n8.json
