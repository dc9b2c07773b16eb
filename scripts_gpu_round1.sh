#!/bin/bash
mkdir -p gpurun_out
echo "== diag" ; timeout 600 python tools/diag_refine_accuracy.py 2>&1 | grep -v "Attention Feature" | tail -8 | tee gpurun_out/diag_accuracy.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError:|relative error|refine_full:" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== bench refine" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-300 ; tail -3 gpurun_out/bench_refine.err
echo "== ncu launch list refine" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn|conv|unfold|fold|l2norm|demote|tc_|cl_|gn_|attention|maxpool|upsample|compose|merge' -c 2000 --csv --log-file gpurun_out/launches_refine.csv python bench.py --workload refine --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_refine.log 2>&1 ; tail -1 gpurun_out/ncu_refine.log | cut -c1-200
