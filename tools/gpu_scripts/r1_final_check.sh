#!/bin/bash
mkdir -p gpurun_out
echo "== bench retrieval"; timeout 200 python bench.py 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | grep -o '"value": [0-9.]*\|"clocks": {[^}]*}\|"e2e": {"value": [0-9.]*' | head -5; grep -i "clocks" gpurun_out/bench_n1.err | head -3
echo "== bench refine"; timeout 120 python bench.py --workload refine --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine_final.json | grep -o '"value": [0-9.]*\|"clocks": {[^}]*}' | head -3
echo "== pytest -m gpu"; timeout 75 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
