#!/bin/bash
# What the driver runs at round end, on one box: GPU tests, smoke, bench (both arms).  Outputs named r02z_*.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/r02z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py > $O/r02z_bench_n1.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -n 3 $O/bench_default.err
python tools/summarize_bench.py $O/r02z_bench_n1.json
timeout 300 python bench.py --impl reference > $O/r02z_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 $O/r02z_bench_reference_arm.json
