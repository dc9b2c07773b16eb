#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/knn_probe.py 2048 2>&1 | tail -14
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/r2_pytest_gpu.log | tail -6
