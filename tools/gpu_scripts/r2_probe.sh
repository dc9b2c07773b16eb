#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/knn_probe.py 2048 2>&1 | tail -12
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/r2_pytest_gpu.log | tail -6
timeout 600 python bench.py > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_full_n1.json'))
print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['retrieval_only'], l.get('cpu_baseline'))
PY
timeout 600 python bench.py --workload retrieval --no-cpu-baseline > gpurun_out/r2_bench_retrieval_n1.json 2>/dev/null
timeout 600 python bench.py --workload retrieval --bank random --no-cpu-baseline > gpurun_out/r2_bench_retrieval_random_n1.json 2>/dev/null
for f in retrieval retrieval_random; do python -c "
import json
l=json.load(open('gpurun_out/r2_bench_${f}_n1.json')); print('$f', l['value'], l['breakdown_ms'], l['e2e']['value'], l.get('knn_stats'), l['roofline']['achieved'])"; done
