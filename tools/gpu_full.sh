#!/bin/bash
# Everything the driver runs at round end, on one box: GPU tests, smoke, bench (both arms).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/summarize_bench.py gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_default.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_default.json
