#!/bin/bash
# ncu evidence for the three kernels (run under gpurun, one GPU): a launch list of the default bench command, then one
# `--set full` capture each of the decode kernel (cfg2), the tcgen05 prefill kernel (cfg5) and the short-sequence
# kernel (cfg4b).  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/run_workload.py cfg2 4 > gpurun_out/plain_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_mma -s 2 -c 1 -f -o gpurun_out/prof_decode \
    python tools/run_workload.py cfg2 4 > gpurun_out/ncu_decode.log 2>&1
echo "decode rc=$?"
python tools/run_workload.py cfg5 3 > gpurun_out/plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o gpurun_out/prof_prefill_cfg5 \
    python tools/run_workload.py cfg5 3 > gpurun_out/ncu_cfg5.log 2>&1
echo "cfg5 rc=$?"
python tools/run_workload.py cfg4b 3 > gpurun_out/plain_cfg4b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prefill_short -s 1 -c 1 -f -o gpurun_out/prof_short_cfg4b \
    python tools/run_workload.py cfg4b 3 > gpurun_out/ncu_cfg4b.log 2>&1
echo "cfg4b rc=$?"
python tools/run_workload.py cfg3 3 > gpurun_out/plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o gpurun_out/prof_prefill_cfg3 \
    python tools/run_workload.py cfg3 3 > gpurun_out/ncu_cfg3.log 2>&1
echo "cfg3 rc=$?"
ls -la gpurun_out/*.ncu-rep
