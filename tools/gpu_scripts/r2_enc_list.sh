#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for t in patch32 patch08; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches_$t.csv python tools/profile_$t.py > /dev/null 2>&1
python - <<PY
import csv
lines=[l for l in open('gpurun_out/r2s2_launches_$t.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum']
seq=[(r['Kernel Name'].split('(')[0].replace('void ','').replace('<unnamed>::','')[:50], float(r['Metric Value'].replace(',',''))/1e6, r['Grid Size']) for r in rows]
print('== $t (last pass)')
# last encoder pass = after the last pad_unfold
idx=[i for i,s in enumerate(seq) if 'wrun' in s[0] or 'cin1' in s[0]]
for n,v,g in seq[idx[-1]-1:]:
    if not n.startswith('native') : print(f"{n:50s} {v:9.4f} {g}")
PY
done
