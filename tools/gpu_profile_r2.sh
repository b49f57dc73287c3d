#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Every ncu pass runs only after the same command exited 0 without ncu.
# 1. launch lists (gpu__time_duration per launch) of every prefill-class workload and the hd-60 decode
# 2. --set full captures: resident-K/V kernel on cfg3 / cfg4a (module layout and dense), the repack kernel, the decode
#    kernel on hd 60 and on cfg2 (refreshes profiles/decode_traffic.json), the tile kernel on the cfg5 shard
mkdir -p gpurun_out
O=gpurun_out
list() {  # name, args...
  local name=$1; shift
  python tools/run_workload.py "$@" > $O/plain_$name.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02s_launches_$name.csv \
      python tools/run_workload.py "$@" > $O/ncu_l_$name.log 2>&1
  echo "launch list $name rc=$?"
}
full() {  # name, kernel regex, skip, args...
  local name=$1 re=$2 skip=$3; shift 3
  python tools/run_workload.py "$@" > $O/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $O/r02s_$name \
      python tools/run_workload.py "$@" > $O/ncu_f_$name.log 2>&1
  echo "full $name rc=$?"
  # the reports are ~22 MB each and gpurun brings back at most 64 MiB: keep the raw page (and the per-instruction
  # source page of the kernels under study) as CSV, drop the report
  ncu -i $O/r02s_$name.ncu-rep --page raw --csv > $O/r02s_ncu_full_${name}_raw.csv 2>/dev/null
  case $name in mid_cfg3|tc_cfg5b2) ncu -i $O/r02s_$name.ncu-rep --page source --csv --print-source sass > $O/r02s_ncu_full_${name}_source.csv 2>/dev/null;; esac
  rm -f $O/r02s_$name.ncu-rep
}
list cfg1t cfg1t 3
list cfg1s cfg1s 3
list cfg1 cfg1 3
list cfg3 cfg3 3
list cfg4a_dense cfg4a 3
list cfg4a_module cfg4a 3 --module-layout
list cfg4b cfg4b 3
list cfg5b2 cfg5b2 3
list cfg2m cfg2m 3
full mid_cfg3 prefill_mid 1 cfg3 3
full mid_cfg4a_module prefill_mid 1 cfg4a 3 --module-layout
full mid_cfg4a_dense prefill_mid 1 cfg4a 3
full repack_cfg4a repack_kernel 1 cfg4a 3
full decode_cfg2m decode_mma 1 cfg2m 3
full decode_cfg2 decode_mma 1 cfg2 3
full tc_cfg5b2 prefill_tc 1 cfg5b2 3
ls -la $O/r02s_*
