#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r2s3_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2s3_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2s3_smoke.log
