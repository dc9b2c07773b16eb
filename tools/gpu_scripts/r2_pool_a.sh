#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "pools_in_epilogue or fused_front or retrieval_backbone or refine_full or unet_backbone or end_to_end or surface" > gpurun_out/r2s3_pytest_i.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s3_pytest_i.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_i.json 2> gpurun_out/r2s3_bench_i.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_i.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_i.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
