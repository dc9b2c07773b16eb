#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in "16384 4 64 64 1 128" "16384 4 64 64 1" "16384 4 32 64 1" "16384 2 64 128 1" "16384 2 64 64 1" "16384 4 32 32 1" "16384 8 32 56 1 64" "16384 8 56 16 1"; do
echo "== $c"; timeout 300 python tools/wp_geo_sweep.py $c 2>&1 | tail -2 | cut -c1-330
done
timeout 600 python tools/wp_layer_times.py 2>&1 | tail -12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_h.json 2> gpurun_out/r2s3_bench_h.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s3_bench_h.err
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_h.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'launches', l['launches_per_step'])"
