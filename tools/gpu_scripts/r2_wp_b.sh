#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in "16384 16 8 16 1" "16384 8 56 16 1" "16384 8 16 16 1" "16384 8 16 32 1" "128 64 16 16 1" "4096 30 8 16 0" "1024 46 16 32 0"; do
echo "== $c"; timeout 300 python tools/wp_geo_sweep.py $c 2>&1 | tail -2
done
