#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/mlp_phases.py attention 2>&1 | head -6
timeout 300 python tools/mlp_phases.py 2>&1 | head -2
timeout 300 python tools/attention_time.py 64 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mlp or encoder or attention" 2>&1 | tail -2
