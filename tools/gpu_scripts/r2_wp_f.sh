#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 2>&1 | grep -v "item [0145]"
RF_WP_SHIFT=1 timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 2>&1 | grep -v "item [0145]"
RF_WP_SHIFT=1 timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 0,1,2,16,1,0 1 2>&1 | grep -v "item [0145]"
RF_WP_SHIFT=1 timeout 900 python -m pytest tests -x -q -m gpu -k "w_pairs" 2>&1 | tail -3
