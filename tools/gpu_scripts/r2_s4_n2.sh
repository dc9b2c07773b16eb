#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02s4_bench_full_n2.json 2> gpurun_out/r02s4_bench_full_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r02s4_bench_full_n2.err
python -c "
import json
l=json.load(open('gpurun_out/r02s4_bench_full_n2.json')); print('n2', l['value'], l['n_gpus'], l['breakdown_ms'], 'e2e', l['e2e']['value'])"
