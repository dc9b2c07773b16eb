#!/bin/bash
# fourth session, step c: GroupNorm statistics (warp per small sample / bulk copies), tests, full-path bench
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
RF_GN_MODE=2 timeout 300 python tools/gn_stats_time.py 2>&1 | tee $O/r02s4_gn_mode2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine or groupnorm or graphs or backbone or decoder or end_to_end or host_pipeline or conv" > $O/r02s4_pytest_c.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s4_pytest_c.log
timeout 900 python bench.py > $O/r02s4_bench_full_c.json 2> $O/r02s4_bench_full_c.err; echo "full rc=$?"; tail -3 $O/r02s4_bench_full_c.err
python -c "
import json
l=json.load(open('$O/r02s4_bench_full_c.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])
for k,v in l['op_breakdown_eager'].items(): print(k, v)"
