#!/bin/bash
# round 2: kNN rework (second pass) + ncu launch list of the full-path step
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "knn or demotion or retrieval or sharding" > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_knn.log
for b in encoded random; do
  timeout 300 python bench.py --workload retrieval --bank $b --no-cpu-baseline > gpurun_out/r2_bench_retrieval_$b.json 2> gpurun_out/r2_bench_retrieval_$b.err; echo "bench $b rc=$?"
  python - <<PY
import json
l=json.load(open('gpurun_out/r2_bench_retrieval_$b.json'))
print('$b', l['value'], l['breakdown_ms'], l['e2e']['value'], l.get('knn_stats'), l['roofline']['achieved'], l['roofline']['frac'])
PY
done
# per-kernel launch list of the full-path step (eager launches; the last step's kernels are summarised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --chunks 16 --refine-batch 16 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_ncu.py launches gpurun_out/r2_launches_full.csv gpurun_out/r2_launches_full.txt; head -40 gpurun_out/r2_launches_full.txt
