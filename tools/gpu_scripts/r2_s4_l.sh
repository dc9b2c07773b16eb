#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/mlp_phases.py attention 2>&1 | head -1
timeout 300 python tools/attention_time.py 64 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or encoder or mlp or refine_full or conv_encoders" 2>&1 | tail -2
