#!/bin/bash
# fourth session: sensitivity of the whole-path throughput to the chunks per step
cd "$GRAFT_REPO_ROOT"
for c in 32 128; do
  timeout 600 python bench.py --no-cpu-baseline --chunks $c --refine-batch $c > gpurun_out/r02s4_bench_full_chunks$c.json 2> gpurun_out/r02s4_bench_full_chunks$c.err; echo "chunks $c rc=$?"
  python -c "
import json
l=json.load(open('gpurun_out/r02s4_bench_full_chunks$c.json')); print($c, 'value', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['clocks']['sm_mhz'])"
done
