#!/bin/bash
# session 2: parity subset + full-path bench
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "shifted_window or backbone or decoder or refine or config3 or surface or encoders" > gpurun_out/r2s2_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s2_pytest_a.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_a.json 2> gpurun_out/r2s2_bench_full_a.err; echo "bench rc=$?"; tail -5 gpurun_out/r2s2_bench_full_a.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2s2_bench_full_a.json'))
print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])
for k,v in l['op_breakdown_eager'].items(): print(k, v)
PY
