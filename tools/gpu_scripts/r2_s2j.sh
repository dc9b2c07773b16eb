#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "single_channel or shifted_window or backbone or decoder or refine or encoders" > gpurun_out/r2s2_pytest_j.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_j.log
for c in 4096,8,32,64,56,0; do
echo "mixed: $(python tools/test_halo_conv.py --case $c 2>&1 | tail -1 | sed 's/geo=.*| halo/| halo/')"
echo "two-pass: $(RF_HALO_FUSED=0 python tools/test_halo_conv.py --case $c 2>&1 | tail -1 | sed 's/geo=.*| halo/| halo/')"
done
python tools/test_halo_conv.py --quick 2>&1 | tail -14
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_bench_full_j.json 2> /dev/null; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_full_j.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
