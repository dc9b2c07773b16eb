#!/bin/bash
# Round-1 (session 3) check: parity tests, re-index kernel timings, headline bench (no CPU leg), launch list.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== reindex timings"; timeout 300 python tools/profile_reindex.py 2>&1 | tee gpurun_out/reindex_times.txt
echo "== bench N=1" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1_nocpu.json | cut -c1-300 ; tail -2 gpurun_out/bench_n1.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn|tc_mlp|pad_unfold|l2norm|RadixSort|tc_linear|demote" -c 400 --csv --log-file gpurun_out/launches_retrieval.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/launches_retrieval.csv gpurun_out/launches_retrieval.txt; head -20 gpurun_out/launches_retrieval.txt
