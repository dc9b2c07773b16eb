#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compose or refine or groupnorm or graphs or backbone or end_to_end or host_pipeline or config3 or surface" > $O/r02s4_pytest_i.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02s4_pytest_i.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --no-cpu-baseline > $O/r02s4_bench_full_i.json 2> $O/r02s4_bench_full_i.err; echo "full rc=$?"
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_i.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'], l['clocks']['sm_mhz'])
for k,v in list(l['op_breakdown_eager'].items()): print(k, v)"
