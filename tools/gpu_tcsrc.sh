#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/run_workload.py cfg5b2 2 > $O/plain_tc.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/tcsrc python tools/run_workload.py cfg5b2 2 > $O/ncu_tc.log 2>&1; echo "rc=$?"
ncu -i $O/tcsrc.ncu-rep --page source --csv --print-source sass > $O/r02w_ncu_full_tc_cfg5b2_source.csv 2>/dev/null
ncu -i $O/tcsrc.ncu-rep --page raw --csv > $O/r02w_ncu_full_tc_cfg5b2_raw.csv 2>/dev/null
rm -f $O/tcsrc.ncu-rep; ls -la $O/r02w*
