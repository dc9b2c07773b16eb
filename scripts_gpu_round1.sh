#!/bin/bash
# Round-1 measurement run: GPU parity tests, the three bench workloads, ncu launch lists and full captures.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
echo "== bench retrieval"; timeout 600 python bench.py 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-250; tail -1 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-250
echo "== bench refine"; timeout 600 python bench.py --workload refine 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-250
echo "== bench stages"; timeout 600 python bench.py --workload stages 2>/dev/null > gpurun_out/stages.json; wc -c gpurun_out/stages.json
echo "== reindex timings"; timeout 300 python tools/profile_reindex.py 2>&1 | tee gpurun_out/reindex_times.txt
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn|tc_mlp|pad_unfold|l2norm|RadixSort|tc_linear|demote" -c 400 --csv --log-file gpurun_out/launches_retrieval.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_refine.csv python bench.py --workload refine --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > /dev/null 2>&1
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates -c 1 -o gpurun_out/knn_cand python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fold_unfold|pad_unfold|compose" -c 16 -o gpurun_out/reindex python tools/profile_reindex.py --once > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"knn_tc_rerank" -c 1 -o gpurun_out/knn_rerank python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | head -40
