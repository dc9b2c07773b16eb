#!/bin/bash
# per-kernel launch list of one full-path step (eager launches); prints the last refine forward
cd "$GRAFT_REPO_ROOT"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > gpurun_out/r2s2_ncu.log 2>&1; echo "ncu rc=$?"
python tools/print_last_step.py gpurun_out/r2s2_launches.csv
