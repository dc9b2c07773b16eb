#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|max abs err|ours-fp64|AssertionError:|relative error|refine_full:" | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== bench N=2" ; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | cut -c1-600 ; tail -5 gpurun_out/bench_n2.err
echo "== bench N=1" ; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-300 ; tail -2 gpurun_out/bench_n1.err
echo "== bench refine N=1" ; timeout 600 python bench.py --workload refine --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_refine.err | tee gpurun_out/bench_refine.json | cut -c1-300 ; tail -2 gpurun_out/bench_refine.err
