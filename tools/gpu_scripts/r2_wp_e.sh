#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "w_pairs or shifted_window or single_channel or stride2" > gpurun_out/r2s3_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s3_pytest_e.log
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 2>&1 | tail -4
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 0,1,2,16,1,0 1 2>&1 | tail -4
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 0 2>&1 | tail -4
timeout 600 python tools/wp_layer_times.py 2>&1 | tail -12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s3_bench_e.json 2> gpurun_out/r2s3_bench_e.err; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2s3_bench_e.json')); print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
