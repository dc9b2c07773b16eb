#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python tools/attention_time.py 64 2>&1 | tee $O/r02s4_attention_time.txt
timeout 300 python tools/mlp_phases.py attention 2>&1 | head -2
timeout 300 python tools/mlp_phases.py 2>&1 | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine or mlp or encoder or graphs or end_to_end or host_pipeline" > $O/r02s4_pytest_h.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02s4_pytest_h.log
timeout 900 python bench.py --no-cpu-baseline > $O/r02s4_bench_full_h.json 2> $O/r02s4_bench_full_h.err; echo "full rc=$?"
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_h.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['clocks'])
for k,v in list(l['op_breakdown_eager'].items())[:5]: print(k, v)"
