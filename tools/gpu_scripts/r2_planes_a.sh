#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "w_pairs or shifted_window or single_channel or stride2 or fused_front or final_decoder or unet_backbone" > gpurun_out/r2s3_pytest_g.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s3_pytest_g.log
for c in "16384 16 8 16 1" "128 64 16 16 1"; do
echo "== $c"; timeout 300 python tools/wp_geo_sweep.py $c 2>&1 | tail -2
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_g.json 2> gpurun_out/r2s3_bench_g.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_g.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_g.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
