#!/bin/bash
# round 2, session 4: tests + bench lines + ncu evidence for profiles/ (one GPU)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > $O/r02s4_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02s4_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02s4_smoke.log
timeout 900 python bench.py > $O/r02s4_bench_full_n1.json 2> $O/r02s4_bench_full_n1.err; echo "full rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02s4_bench_reference_arm.json 2> $O/r02s4_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py --workload retrieval --no-cpu-baseline > $O/r02s4_bench_retrieval_encoded.json 2>/dev/null; echo "retr rc=$?"
timeout 600 python bench.py --workload surface --no-cpu-baseline > $O/r02s4_bench_surface_n1.json 2>/dev/null; echo "surface rc=$?"
timeout 900 python bench.py --workload stages > $O/r02s4_stage_rooflines.json 2>/dev/null; echo "stages rc=$?"
timeout 300 python tools/attention_time.py 64 2>&1 | tee $O/r02s4_attention_time.txt
timeout 300 python tools/mlp_phases.py attention > $O/r02s4_mlp_phases_attention.txt 2>&1
# launch list of one eager full-path step
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02s4_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/print_last_step.py $O/r02s4_launches_full.csv > $O/r02s4_full_path_last_step.txt; tail -1 $O/r02s4_full_path_last_step.txt
# --set full captures: the GroupNorm statistics kernels of this session, the attention's kernels at 64 chunks
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_gn_partial_warp_kernel -s 4 -c 1 -f -o $O/r02s4_gn_warp_16ch_8cube \
  python tools/gn_stats_time.py > /dev/null 2>&1; echo "ncu gn warp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_gn_partial_bulk_kernel -c 1 -f -o $O/r02s4_gn_bulk_8ch_16cube \
  python tools/gn_stats_time.py > /dev/null 2>&1; echo "ncu gn bulk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:unfold_e2_patch8_kernel|attention_epilogue_kernel|tc_mlp_kernel" -s 4 -c 4 -f -o $O/r02s4_attention_64chunks \
  python tools/attention_time.py 64 --once > /dev/null 2>&1; echo "ncu attention rc=$?"
ls -la $O/r02s4_* | tail -30
