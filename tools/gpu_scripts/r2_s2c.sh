#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "single_channel or shifted_window" > gpurun_out/r2s2_pytest_c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s2_pytest_c.log
timeout 900 python -m pytest tests -x -q -m gpu -k "backbone or decoder or refine or config3 or surface or encoders or hot_path" > gpurun_out/r2s2_pytest_c2.log 2>&1; echo "pytest2 rc=$?"; tail -15 gpurun_out/r2s2_pytest_c2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_c.json 2> gpurun_out/r2s2_bench_full_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s2_bench_full_c.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2s2_bench_full_c.json'))
print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])
for k,v in l['op_breakdown_eager'].items(): print(k, v)
PY
timeout 600 python bench.py --workload surface --no-cpu-baseline > gpurun_out/r2s2_bench_surface_c.json 2> /dev/null; python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_surface_c.json')); print('surface', l['value'], l['breakdown_ms'])"
