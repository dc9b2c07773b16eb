#!/bin/bash
# fourth session: BASELINE configs[2] as specified - bank in 4 row shards, full path on 4 GPUs
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --bank-shards 4 --no-cpu-baseline > gpurun_out/r02s4_bench_full_n4_shards4.json 2> gpurun_out/r02s4_bench_full_n4_shards4.err; echo "n4 rc=$?"; tail -3 gpurun_out/r02s4_bench_full_n4_shards4.err
python -c "
import json
l=json.load(open('gpurun_out/r02s4_bench_full_n4_shards4.json')); print('n4', l['value'], l['n_gpus'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
