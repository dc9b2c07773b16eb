#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=8 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError|relative error|refine_full:|s call|s setup" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== bench retrieval m4" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --knn-method 4 2> gpurun_out/bench_m4.err | tee gpurun_out/bench_m4.json | cut -c1-200; tail -2 gpurun_out/bench_m4.err; grep -o '"breakdown_ms.*' gpurun_out/bench_m4.json | cut -c1-500
echo "== bench refine (graph)" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-300 ; tail -3 gpurun_out/bench_refine.err
