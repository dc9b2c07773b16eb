#!/bin/bash
# fourth session: ncu launch list of one eager full-path step
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02s4_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/print_last_step.py $O/r02s4_launches_full.csv > $O/r02s4_full_path_last_step.txt; tail -1 $O/r02s4_full_path_last_step.txt
