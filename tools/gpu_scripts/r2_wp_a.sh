#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "w_pairs" > gpurun_out/r2s3_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2s3_pytest_a.log
timeout 600 python tools/wp_layer_times.py 2>&1 | tail -12
