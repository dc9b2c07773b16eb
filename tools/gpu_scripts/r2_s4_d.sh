#!/bin/bash
# fourth session, step d: fused MLP chain - batched input loads, incremental MMA issue loop
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python tools/mlp_phases.py attention 2>&1 | tee $O/r02s4_mlp_phases_attention.txt
timeout 300 python tools/mlp_phases.py 2>&1 | tee $O/r02s4_mlp_phases_patch04.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine or mlp or encoder or graphs or end_to_end or host_pipeline" > $O/r02s4_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s4_pytest_d.log
timeout 900 python bench.py > $O/r02s4_bench_full_d.json 2> $O/r02s4_bench_full_d.err; echo "full rc=$?"; tail -3 $O/r02s4_bench_full_d.err
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_d.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])
for k,v in l['op_breakdown_eager'].items(): print(k, v)
print(l['retrieval_only'])"
