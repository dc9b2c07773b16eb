#!/bin/bash
# final state of the fourth session: tests, smoke, bench line (N = 1)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > $O/r02s4_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s4_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02s4_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02s4_smoke.log
timeout 900 python bench.py > $O/r02s4_bench_full_n1_final.json 2> $O/r02s4_bench_full_n1_final.err; echo "full rc=$?"
timeout 900 python bench.py --workload stages > $O/r02s4_stage_rooflines.json 2>/dev/null; echo "stages rc=$?"
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_n1_final.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'], l['clocks']); r=l['roofline']; print({k:r[k] for k in ('achieved','frac','ms_per_step','share_of_step','traffic')})
s=json.load(open('$O/r02s4_stage_rooflines.json'))['stages']
for k,v in s.items(): print(k, round(v['ms'],3))"
