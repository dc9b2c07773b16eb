#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for b in encoded random; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn --csv --log-file gpurun_out/r2s2_knn_$b.csv python bench.py --workload retrieval --bank $b --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv, collections
lines=[l for l in open('gpurun_out/r2s2_knn_$b.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
seq=[(r['Kernel Name'].split('(')[0][-45:], float(r['Metric Value'].replace(',',''))/1e6, r['Grid Size']) for r in rows]
print('$b')
for n,v,g in seq[-40:]:
    if v>0.05: print('  ',n,round(v,3),g)
PY
done
