#!/bin/bash
# ncu --set full of the two backward kernels on the training-step workload, final code (after the plain run exited 0)
mkdir -p gpurun_out
O=gpurun_out
python tools/run_backward.py > $O/plain_bwd.log 2>&1 || exit 1
for k in attn_bwd_dq attn_bwd_dkv; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/r02z_$k python tools/run_backward.py > $O/ncu_$k.log 2>&1; echo "full $k rc=$?"
  ncu -i $O/r02z_$k.ncu-rep --page raw --csv > $O/r02z_ncu_full_${k}_raw.csv 2>/dev/null
  rm -f $O/r02z_$k.ncu-rep
done
ls -la $O/r02z_ncu*
