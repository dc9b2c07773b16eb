#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (subset)" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "tc_linear or encoder or knn_bit" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_gpu_subset.log
echo "== bench retrieval" ; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_retrieval.err | tee gpurun_out/bench_retrieval.json ; tail -2 gpurun_out/bench_retrieval.err
echo "== ncu launch list (our kernels only)" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn|conv_igemm|pad_unfold|l2norm|demote|tc_linear|tc_weight|merge' -c 200 --csv --log-file gpurun_out/launches_retrieval.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --bank random > gpurun_out/ncu_bench.log 2>&1 ; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
