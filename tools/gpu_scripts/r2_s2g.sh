#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "single_channel or shifted_window or backbone or decoder or refine or encoders or surface" > gpurun_out/r2s2_pytest_g.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_g.log
for c in 16,64,16,0,16,0 64,64,16,0,16,0 4096,16,8,0,16,0; do
echo "case $c: $(python tools/test_halo_conv.py --case $c 2>&1 | tail -1 )"
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_g.json 2> gpurun_out/r2s2_bench_full_g.err; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_full_g.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
timeout 600 python bench.py --workload surface --no-cpu-baseline > gpurun_out/r2s2_bench_surface_g.json 2> /dev/null; python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_surface_g.json')); print('surface', l['value'], l['breakdown_ms'])"
