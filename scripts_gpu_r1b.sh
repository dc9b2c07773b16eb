#!/bin/bash
# Round-1 (session 3) check of the rewritten re-indexing kernels: parity tests, event timings, ncu captures.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== reindex timings"; timeout 300 python tools/profile_reindex.py 2>&1 | tee gpurun_out/reindex_times.txt
echo "== ncu reindex"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fold_unfold|pad_unfold|compose" -c 16 -o gpurun_out/reindex python tools/profile_reindex.py --once > /dev/null 2>&1; ls -la gpurun_out/reindex.ncu-rep
echo "== bench stages"; timeout 600 python bench.py --workload stages 2>gpurun_out/stages.err > gpurun_out/stages.json; wc -c gpurun_out/stages.json; grep -c stages gpurun_out/stages.err
