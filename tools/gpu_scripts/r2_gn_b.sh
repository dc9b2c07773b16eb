#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/gn_epi_time.py 2>&1 | tail -10
timeout 900 python -m pytest tests -x -q -m gpu -k "groupnorm_in_epilogue" 2>&1 | tail -3
