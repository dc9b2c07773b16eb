#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=6 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError|relative error|refine_full:|s call|s setup|Error" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== ncu full: halo conv 96->56 @ 8^3 x 2048 patches"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv3d_halo -c 1 -o gpurun_out/halo_conv_96_56 python tools/test_halo_conv.py --case 2048,8,32,64,56,0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:tc_conv3d_halo -c 1 -o gpurun_out/halo_conv_dec16 python tools/test_halo_conv.py --case 8,64,16,0,16,0 > /dev/null 2>&1
echo "== ncu launch list (refine)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_refine.csv python bench.py --workload refine --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_refine.log 2>&1; tail -1 gpurun_out/ncu_refine.log | cut -c1-200
echo "== bench refine (graph)" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-300 ; tail -2 gpurun_out/bench_refine.err
ls -la gpurun_out
