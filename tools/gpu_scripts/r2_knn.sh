#!/bin/bash
# round 2: kNN rework - parity tests + retrieval-only bench on the encoded and the isotropic bank
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "knn or demotion or retrieval or sharding" > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_knn.log
for b in encoded random; do
  timeout 300 python bench.py --workload retrieval --bank $b --no-cpu-baseline > gpurun_out/r2_bench_retrieval_$b.json 2> gpurun_out/r2_bench_retrieval_$b.err; echo "bench $b rc=$?"
  python - <<PY
import json
l=json.load(open('gpurun_out/r2_bench_retrieval_$b.json'))
print('$b', l['value'], l['breakdown_ms'], l['e2e']['value'], l.get('knn_stats'), l['roofline']['achieved'], l['roofline']['frac'])
PY
done
