#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_bwd.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/t_bwd.log
timeout 200 python tools/run_backward.py --time | grep forward
timeout 200 python tools/run_backward.py 4 4096 32 8 128 -1 --time | grep forward
timeout 200 python tools/run_backward.py 64 196 16 8 72 -1 --time | grep forward
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02z_launches_train_step.csv python tools/run_backward.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/r02z_launches_train_step.csv') if not l.startswith('=='))]
for r in rows[-4:]:
    print(r['Kernel Name'][:50], r['Metric Name'], r['Metric Value'])
PY
