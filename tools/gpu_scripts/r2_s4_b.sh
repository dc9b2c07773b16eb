#!/bin/bash
# fourth session, step b: GroupNorm statistics variants
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for m in 1 2 3; do RF_GN_MODE=$m timeout 300 python tools/gn_stats_time.py; done 2>&1 | tee $O/r02s4_gn_modes.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "groupnorm or backbone or refine" 2>&1 | tail -3
