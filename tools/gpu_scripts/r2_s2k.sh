#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "host_pipeline" > gpurun_out/r2s2_pytest_k.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_k.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_k.json 2> gpurun_out/r2s2_bench_full_k.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s2_bench_full_k.err
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_full_k.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e'])"
