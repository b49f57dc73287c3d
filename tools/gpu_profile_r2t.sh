#!/bin/bash
# Last session of round 2: everything the driver runs at round end + ncu of the two kernels that changed
# (repack_chunk_kernel, decode_mma_kernel<64> with the 12-stage ring).  ncu passes only after the plain command exited 0.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/r02t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py > $O/r02t_bench_n1.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -n 3 $O/bench_default.err
python tools/summarize_bench.py $O/r02t_bench_n1.json
timeout 300 python bench.py --impl reference > $O/r02t_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 $O/r02t_bench_reference_arm.json
full() {  # name, kernel regex, skip, args...
  local name=$1 re=$2 skip=$3; shift 3
  python tools/run_workload.py "$@" > $O/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $O/r02t_$name \
      python tools/run_workload.py "$@" > $O/ncu_f_$name.log 2>&1
  echo "full $name rc=$?"
  ncu -i $O/r02t_$name.ncu-rep --page raw --csv > $O/r02t_ncu_full_${name}_raw.csv 2>/dev/null
  rm -f $O/r02t_$name.ncu-rep
}
python tools/run_workload.py cfg4a 3 > $O/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02t_launches_cfg4a_dense.csv python tools/run_workload.py cfg4a 3 > $O/ncu_l.log 2>&1; echo "launch list rc=$?"
full repack_cfg4a repack_chunk 1 cfg4a 3
full decode_cfg2m decode_mma 1 cfg2m 3
ls -la $O/r02t_*
