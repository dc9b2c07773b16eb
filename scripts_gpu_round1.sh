#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError:|relative error|refine_full:" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== bench refine" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-300 ; tail -3 gpurun_out/bench_refine.err
echo "== bench retrieval" ; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_retrieval.err | tee gpurun_out/bench_retrieval.json | cut -c1-300; tail -2 gpurun_out/bench_retrieval.err
