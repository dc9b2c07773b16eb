#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "groupnorm_in_epilogue" > gpurun_out/r2s3_pytest_k.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2s3_pytest_k.log
timeout 900 python -m pytest tests -x -q -m gpu -k "retrieval_backbone or refine_full or unet_backbone or end_to_end or surface or final_decoder or pools" > gpurun_out/r2s3_pytest_k2.log 2>&1; echo "pytest2 rc=$?"; tail -5 gpurun_out/r2s3_pytest_k2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_k.json 2> gpurun_out/r2s3_bench_k.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_k.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_k.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
