#!/bin/bash
# ncu launch list of the headline bench command (after the same command exited 0 without ncu)
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 3 --warmup 3 --no-extra > $O/plain_bench.json 2>$O/plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02z_launches_bench_decode.csv python bench.py --steps 3 --warmup 3 --no-extra > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/r02z_launches_bench_decode.csv') if not l.startswith('=='))]
c=collections.defaultdict(list)
for r in rows: c[r['Kernel Name'][:60]].append(float(r['Metric Value']))
for k,v in sorted(c.items(), key=lambda kv:-sum(kv[1]))[:8]: print(k, len(v), round(sum(v)/len(v)/1000,1),'us mean')
PY
