#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine or graphs or end_to_end or host_pipeline" > $O/r02s4_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s4_pytest_e.log
timeout 900 python bench.py --no-cpu-baseline > $O/r02s4_bench_full_e.json 2> $O/r02s4_bench_full_e.err; echo "full rc=$?"; tail -3 $O/r02s4_bench_full_e.err
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_e.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])
for k,v in list(l['op_breakdown_eager'].items())[:6]: print(k, v)"
