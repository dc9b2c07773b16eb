#!/bin/bash
# round 2, first GPU pass: parity tests, the full-path bench, the CPU arm
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_full_n1.json | head -c 6000
tail -5 gpurun_out/r2_bench_full_n1.err
python bench.py --refine-batch 8 --no-cpu-baseline > gpurun_out/r2_bench_full_rb8.json 2>/dev/null
python bench.py --refine-batch 32 --no-cpu-baseline > gpurun_out/r2_bench_full_rb32.json 2>/dev/null
python bench.py --refine-batch 64 --no-cpu-baseline > gpurun_out/r2_bench_full_rb64.json 2>/dev/null
for f in rb8 rb32 rb64; do python -c "
import json,sys
l=json.load(open('gpurun_out/r2_bench_full_$f.json')); print('$f', l['value'], l['breakdown_ms'], l['e2e']['value'])"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref.json
