#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25 | cut -c1-400 | tee gpurun_out/pytest_gpu.log
echo "== bench refine"; timeout 600 python bench.py --workload refine --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine_nocpu.json | cut -c1-700
