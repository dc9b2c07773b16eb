#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s > gpurun_out/r2_pytest_bwd.log 2>&1; echo "bwd rc=$?"
tail -25 gpurun_out/r2_pytest_bwd.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|config 3|surface config|refine_full" gpurun_out/r2_pytest_gpu.log | tail -12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_full_n1.json'))
print('full', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'], l['retrieval_only'])
print({k:(v['ms'],v.get('tflops'),v.get('gbs')) for k,v in list(l['op_breakdown_eager'].items())[:8]})
PY
timeout 900 python bench.py --workload surface --no-cpu-baseline > gpurun_out/r2_bench_surface_n1.json 2> gpurun_out/r2_bench_surface_n1.err; echo "surface rc=$?"; tail -3 gpurun_out/r2_bench_surface_n1.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_surface_n1.json'))
print('surface', l['value'], l['breakdown_ms'], 'e2e', l['e2e']['value'])
print({k:(v['ms'],v.get('tflops'),v.get('gbs')) for k,v in list(l['op_breakdown_eager'].items())[:8]})
PY
timeout 900 python bench.py --workload sweep --steps 3 > gpurun_out/r2_bench_sweep_n1.json 2> gpurun_out/r2_bench_sweep_n1.err; echo "sweep rc=$?"; grep sweep gpurun_out/r2_bench_sweep_n1.err | cut -c1-400
