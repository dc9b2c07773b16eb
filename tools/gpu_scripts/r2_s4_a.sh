#!/bin/bash
# fourth session, step a: float4 GroupNorm statistics, attention on un-folded patches with channels-last store
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine or groupnorm or graphs or backbone or decoder or end_to_end or host_pipeline" > $O/r02s4_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02s4_pytest_a.log
timeout 900 python bench.py > $O/r02s4_bench_full_a.json 2> $O/r02s4_bench_full_a.err; echo "full rc=$?"; tail -3 $O/r02s4_bench_full_a.err
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_a.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])
for k,v in l['op_breakdown_eager'].items(): print(k, v)"
