#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 0 2>&1 | tail -7
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 2>&1 | tail -7
timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 0,1,2,16,1,0 1 2>&1 | tail -7
timeout 100 python tools/wp_phases.py 16384 8 16 16 1 0 2>&1 | tail -7
timeout 100 python tools/wp_phases.py 16384 8 16 16 1 1 2>&1 | tail -7
