#!/bin/bash
# round 2: whole GPU suite + the bench lines of every BASELINE config
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|config 3|surface config|refine_full" gpurun_out/r2_pytest_gpu.log | tail -12
timeout 600 python bench.py > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_full_n1.json'))
print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'cpu', l.get('cpu_baseline'))
print({k:(v['ms'],v.get('tflops'),v.get('gbs')) for k,v in list(l['op_breakdown_eager'].items())[:8]})
print(l['roofline']['achieved'], l['roofline']['frac'], l['roofline']['share_of_step'])
PY
timeout 900 python bench.py --workload surface > gpurun_out/r2_bench_surface_n1.json 2> gpurun_out/r2_bench_surface_n1.err; echo "surface rc=$?"; tail -3 gpurun_out/r2_bench_surface_n1.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_surface_n1.json'))
print('surface', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], 'cpu', l.get('cpu_baseline'))
print({k:(v['ms'],v.get('tflops'),v.get('gbs')) for k,v in list(l['op_breakdown_eager'].items())[:8]})
PY
timeout 900 python bench.py --workload sweep --steps 3 > gpurun_out/r2_bench_sweep_n1.json 2> gpurun_out/r2_bench_sweep_n1.err; echo "sweep rc=$?"; tail -6 gpurun_out/r2_bench_sweep_n1.err
