#!/bin/bash
# last check of the final build: the whole GPU suite and the smoke call
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02s4_pytest_gpu_verify.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s4_pytest_gpu_verify.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
