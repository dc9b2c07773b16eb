#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2s2_bench_full_n2.json 2> gpurun_out/r2s2_bench_full_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2s2_bench_full_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2s2_bench_full_n2.json'):
    if l.startswith('{'):
        l=json.loads(l); print('full n2', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['config'].get('bank'), l.get('retrieval_only'))
PY
