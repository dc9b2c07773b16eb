#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches_surface.csv \
  python bench.py --workload surface --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > gpurun_out/r2s2_ncu_surface.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/r2s2_launches_surface.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum']
seq=[(r['Kernel Name'].split('(')[0].replace('void ','').replace('<unnamed>::','')[:50], float(r['Metric Value'].replace(',',''))/1e6, r['Grid Size']) for r in rows]
idx=[i for i,s in enumerate(seq) if 'compose_kernel' in s[0]]
last=idx[-1]
# encode part: from the previous pad_unfold before last compose
st=max(i for i in range(last) if 'pad_unfold' in seq[i][0] or 'occup' in seq[i][0].lower())
print('--- encode..compose of the last step')
for n,v,g in seq[st-3:last+1]:
    if not n.startswith('native') and not n.startswith('cub'): print(f"{n:50s} {v:9.4f} {g}")
agg=collections.OrderedDict()
for n,v,g in seq[last:]:
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
print('--- refine part by kernel')
for n,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{n:50s} {c:5d} {v:9.3f}")
PY
