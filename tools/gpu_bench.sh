#!/bin/bash
# Bench + ncu launch list + one full ncu capture of the decode kernel (run under gpurun).
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_split -s 3 -c 2 -f -o gpurun_out/prof_decode \
    python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
