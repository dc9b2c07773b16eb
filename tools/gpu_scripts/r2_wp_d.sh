#!/bin/bash
cd "$GRAFT_REPO_ROOT"
RF_HALO_NB=8 timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 2>&1 | tail -4
RF_HALO_NB=8 timeout 100 python tools/wp_phases.py 16384 16 8 16 1 1 0,1,2,16,1,0 1 2>&1 | tail -4
RF_HALO_NB=8 timeout 100 python tools/wp_phases.py 16384 16 8 16 1 0 2>&1 | tail -4
RF_HALO_NB=8 timeout 600 python tools/wp_layer_times.py 2>&1 | tail -12
