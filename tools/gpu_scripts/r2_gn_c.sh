#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/gn_epi_time.py 2>&1 | grep "N="
timeout 900 python -m pytest tests -x -q -m gpu -k "groupnorm_in_epilogue or retrieval_backbone or refine_full" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_l.json 2> gpurun_out/r2s3_bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_l.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_l.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
