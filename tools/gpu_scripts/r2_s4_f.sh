#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python bench.py --no-cpu-baseline --steps 20 > $O/r02s4_bench_full_steps20.json 2> /dev/null; echo "rc=$?"
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_steps20.json')); print('full', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], l['e2e'].get('sync_value'), l['clocks'])"
