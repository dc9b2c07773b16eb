#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (attention + end-to-end)" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "attention or refine or surface or end_to_end or metrics or ntxent or sobel" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu_attn.log
echo "== bench refine"; timeout 600 python bench.py --workload refine 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-250
echo "== bench stages"; timeout 600 python bench.py --workload stages 2>/dev/null > gpurun_out/stages.json; wc -c gpurun_out/stages.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"attention_epilogue" -c 2 python bench.py --workload refine --steps 1 --warmup 0 --no-cpu-baseline --no-cuda-graph 2>/dev/null | grep -B1 -A4 "dram__bytes_read" | head -16
