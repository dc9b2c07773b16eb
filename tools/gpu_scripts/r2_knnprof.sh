#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn --csv --log-file gpurun_out/r2_knn_probe.csv python tools/knn_probe.py 2048 > gpurun_out/r2_knn_probe.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/r2_knn_probe.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
for r in rows[:200]:
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v*={'ns':1e-3,'nsecond':1e-3,'us':1,'usecond':1,'ms':1e3,'msecond':1e3}.get(u,1)
    print(r['Kernel Name'][:50], r['Grid Size'], f"{v:.1f} us")
PY
