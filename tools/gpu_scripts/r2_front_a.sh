#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "fused_front or retrieval_backbone or refine_full or w_pairs" > gpurun_out/r2s3_pytest_f.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s3_pytest_f.log
timeout 300 python tools/front_time.py 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_f.json 2> gpurun_out/r2s3_bench_f.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_f.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_f.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
