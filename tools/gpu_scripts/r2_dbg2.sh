#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in "" 1; do
RF_BENCH_NO_CLOCKS=$v timeout 300 python bench.py --workload retrieval --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('noclocks=[$v]', l['value'], l['breakdown_ms'], l['clocks'] and l['clocks']['samples'])"
done
RF_BENCH_NO_CLOCKS=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('full noclocks', l['value'], l['breakdown_ms'])"
