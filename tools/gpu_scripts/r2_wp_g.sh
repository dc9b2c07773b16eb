#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in "16384 8 32 56 1 64" "16384 4 64 64 1 128" "16384 4 64 64 1" "16384 8 16 32 1"; do
echo "== $c"; timeout 300 python tools/wp_geo_sweep.py $c 2>&1 | tail -2
done
