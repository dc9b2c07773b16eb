#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=6 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError|relative error|refine_full:|s call|s setup|Error" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== bench refine (graph)" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-400 ; tail -3 gpurun_out/bench_refine.err
echo "== ncu launch list (refine)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_refine.csv python bench.py --workload refine --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_refine.log 2>&1; tail -2 gpurun_out/ncu_refine.log | cut -c1-200
