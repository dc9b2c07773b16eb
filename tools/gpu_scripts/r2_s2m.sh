#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "stride2 or encoders or surface" > gpurun_out/r2s3_pytest_j.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2s3_pytest_j.log
bash tools/gpu_scripts/r2_enc_list.sh 2>&1 | grep -v "^at::"
timeout 600 python bench.py --workload surface --no-cpu-baseline > /tmp/s.json 2> /dev/null; python -c "
import json
l=json.load(open('/tmp/s.json')); print('surface', l['value'], l['breakdown_ms'])"
