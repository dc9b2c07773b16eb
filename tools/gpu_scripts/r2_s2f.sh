#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s2_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_all.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_f.json 2> gpurun_out/r2s2_bench_full_f.err; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_full_f.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
timeout 600 python bench.py --workload surface --no-cpu-baseline > gpurun_out/r2s2_bench_surface_f.json 2> /dev/null; python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_surface_f.json')); print('surface', l['value'], l['breakdown_ms'])"
timeout 900 python bench.py --workload stages > gpurun_out/r2s2_stages.json 2> gpurun_out/r2s2_stages.err; echo "stages rc=$?"; python -c "
import json
for l in open('gpurun_out/r2s2_stages.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for k,v in (d.get('stages') or d).items() if isinstance(d.get('stages') or d, dict) else []: print(k, v if not isinstance(v, dict) else {a:b for a,b in v.items() if a in ('ms','frac','achieved','unit')})
"
