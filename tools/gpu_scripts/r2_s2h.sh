#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "attention or mlp or refine_full" > gpurun_out/r2s2_pytest_h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_h.log
for wide in 1 0; do
RF_MLP_WIDE=$wide timeout 600 python bench.py --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('wide=$wide full', l['value'], l['breakdown_ms']['refine'], l['op_breakdown_eager']['rf_attention_fuse_fwd'])"
done
