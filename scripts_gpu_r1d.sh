#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (kNN + end-to-end)" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "knn or end_to_end or pipeline or demotion or interface" --durations=5 2>&1 | tail -14 | cut -c1-300 | tee gpurun_out/pytest_gpu_knn.log
echo "== bench N=1" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1_nocpu.json | cut -c1-200 ; grep -o '"breakdown_ms": {[^}]*}' gpurun_out/bench_n1_nocpu.json
echo "== ncu attention kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_mlp_kernel|attention_epilogue" -c 3 -o gpurun_out/attention python bench.py --workload refine --steps 1 --warmup 0 --no-cpu-baseline --no-cuda-graph > /dev/null 2>&1; ls -la gpurun_out/attention.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_tc_rerank" -c 3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -A3 "gpu__time_duration" | head -12
