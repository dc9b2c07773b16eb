#!/bin/bash
# GPU pass: parity tests, smoke, bench (both workloads), ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log
for m in 0 3; do
echo "== bench retrieval method $m" ; timeout 600 python bench.py --steps 5 --warmup 3 --knn-method $m 2> gpurun_out/bench_retrieval_m$m.err | tee gpurun_out/bench_retrieval_m$m.json ; tail -3 gpurun_out/bench_retrieval_m$m.err
done
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_retrieval.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full (knn_tc_candidates)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates -c 1 -o gpurun_out/prof_knn_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
