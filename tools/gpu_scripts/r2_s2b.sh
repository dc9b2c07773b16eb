#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "shifted_window" > gpurun_out/r2s2_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2_pytest_a.log
bash tools/gpu_scripts/r2_launchlist.sh | grep -v "^native\|^cub" | tail -75
