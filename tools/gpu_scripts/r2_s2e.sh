#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in 4096,8,32,64,56,0 4096,16,8,0,16,0 4096,8,56,0,16,0 4096,4,64,128,64,0 16,64,16,0,16,0; do
echo "case $c: $(python tools/test_halo_conv.py --case $c 2>&1 | tail -1 | sed 's/geo=.*| halo/| halo/')"
done
