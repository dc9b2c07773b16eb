#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or refine" > $O/r02s4_pytest_g.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02s4_pytest_g.log
