#!/bin/bash
# fourth session: BASELINE configs[2] as specified - bank in default sharding, full path on 8 GPUs
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02s4_bench_full_n8.json 2> gpurun_out/r02s4_bench_full_n8.err; echo "n4 rc=$?"; tail -3 gpurun_out/r02s4_bench_full_n8.err
python -c "
import json
l=json.load(open('gpurun_out/r02s4_bench_full_n8.json')); print('n4', l['value'], l['n_gpus'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
