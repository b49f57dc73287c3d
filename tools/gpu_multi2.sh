#!/bin/bash
# 2-GPU box: the multi-GPU tests (NCCL world 2 + one process on two devices), the extended smoke, bench at N = 2
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider > $O/r02v_pytest_gpu_multi_2gpu.log 2>&1; echo "multi tests rc=$?"; tail -n 6 $O/r02v_pytest_gpu_multi_2gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 5 > $O/r02v_bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -n 3 $O/bench_n2.err
python tools/summarize_bench.py $O/r02v_bench_n2.json
