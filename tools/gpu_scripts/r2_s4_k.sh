#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python tools/mlp_phases.py attention 2>&1 | head -3
timeout 300 python tools/mlp_phases.py 2>&1 | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not knn and not fold and not patcher and not sweep" > $O/r02s4_pytest_k.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02s4_pytest_k.log
timeout 900 python bench.py --no-cpu-baseline > $O/r02s4_bench_full_k.json 2> $O/r02s4_bench_full_k.err; echo "full rc=$?"
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_k.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['clocks']['sm_mhz'])
for k,v in list(l['op_breakdown_eager'].items())[:6]: print(k, v)"
