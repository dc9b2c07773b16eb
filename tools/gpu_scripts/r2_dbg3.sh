#!/bin/bash
cd "$GRAFT_REPO_ROOT"
b() { timeout 300 python bench.py --workload retrieval --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null; python -c "
import json
l=json.load(open('/tmp/b.json')); print('$1', round(l['value']), l['breakdown_ms']['knn'])"; }
b fresh
timeout 600 python -m pytest tests -x -q -m gpu -k "knn" > /dev/null 2>&1; echo "pytest knn rc=$?"
b after_knn_tests
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
b after_all_tests
python tools/knn_probe2.py 2>&1 | tail -5
