#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu -k "knn or retrieval or demotion or hot_path" > gpurun_out/r2s2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2_pytest_knn.log
for b in encoded random; do
timeout 600 python bench.py --workload retrieval --bank $b --no-cpu-baseline > gpurun_out/r2s2_bench_retrieval_$b.json 2>/dev/null
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_retrieval_$b.json')); print('$b', l['value'], l['breakdown_ms'], l['e2e']['value'], l.get('knn_stats'), l['roofline']['achieved'], l['roofline']['frac'])"
done
timeout 900 python bench.py --workload sweep --steps 2 --warmup 1 > gpurun_out/r2s2_bench_sweep.json 2> gpurun_out/r2s2_bench_sweep.err; echo "sweep rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2s2_bench_sweep.json'))
for s in l['sweep']: print({k:s[k] for k in ('k','knn_bulk_ms','knn_bulk_frac_of_bf16_peak','knn_64chunks_ms','knn_stats','attention_fuse_ms_batch8')})"
