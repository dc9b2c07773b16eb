#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "single_channel or shifted_window or backbone or decoder or refine or encoders" > gpurun_out/r2s2_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_d.log
for res in 0 1; do
RF_HALO_RES=$res timeout 600 python bench.py --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('res=$res full', l['value'], l['breakdown_ms']['refine'])"
done
bash tools/gpu_scripts/r2_launchlist.sh | grep "tc_conv3d_halo_kernel\|wrun\|total"
