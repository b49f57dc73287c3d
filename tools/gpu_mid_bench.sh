#!/bin/bash
# timing of the <= 256-key workloads with the resident-K/V kernel on and off
mkdir -p gpurun_out
for w in cfg3 cfg4a; do
  for lay in "" "--module-layout"; do
    for mid in 1 0; do
      echo "== $w $lay MID=$mid"; VATS_PREFILL_MID=$mid timeout 120 python tools/run_workload.py $w 3 --time --flush $lay 2>&1 | tail -2
    done
  done
done
for w in cfg1t cfg1s; do
  for mid in 1 0; do echo "== $w MID=$mid"; VATS_PREFILL_MID=$mid timeout 120 python tools/run_workload.py $w 3 --time --flush 2>&1 | tail -2; done
done
for w in cfg3 cfg4a; do echo "== $w module-layout bounded"; timeout 120 python tools/run_workload.py $w 3 --time --flush --module-layout --bound 2>&1 | tail -2; done
for w in cfg5b2; do for b in "" "--bound"; do echo "== $w $b"; timeout 120 python tools/run_workload.py $w 3 --time $b 2>&1 | tail -2; done; done
