#!/bin/bash
# round 2, session 2: bench lines + ncu evidence for profiles/ (one GPU)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python bench.py > $O/r02_bench_full_n1.json 2> $O/r02_bench_full_n1.err; echo "full rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py --workload retrieval --no-cpu-baseline > $O/r02_bench_retrieval_encoded.json 2>/dev/null; echo "retr rc=$?"
timeout 600 python bench.py --workload retrieval --bank random --no-cpu-baseline > $O/r02_bench_retrieval_random.json 2>/dev/null; echo "retr random rc=$?"
timeout 900 python bench.py --workload sweep --steps 3 --warmup 2 > $O/r02_bench_sweep_n1.json 2>/dev/null; echo "sweep rc=$?"
timeout 600 python bench.py --workload surface > $O/r02_bench_surface_n1.json 2>/dev/null; echo "surface rc=$?"
timeout 900 python bench.py --workload stages > $O/r02_stage_rooflines.json 2>/dev/null; echo "stages rc=$?"
# launch list + DRAM traffic of the dominant kernel over one eager full-path step
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_conv3d_halo_kernel \
  --csv --log-file $O/r02_traffic_conv.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu traffic rc=$?"
# --set full captures
for c in 4096,8,32,64,56,0 4096,16,8,0,16,0 4096,4,64,128,64,0 16,64,16,0,16,0; do
  n=$(echo $c | tr ',' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3d_halo_kernel -c 1 -f -o $O/r02_halo_$n \
    python tools/test_halo_conv.py --case $c > /dev/null 2>&1; echo "ncu full $c rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_norm_split_halo_kernel -c 1 -f -o $O/r02_split_4096_8_32_64 \
  python tools/test_halo_conv.py --case 4096,8,32,64,56,0 > /dev/null 2>&1; echo "ncu split rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates_kernel -s 2 -c 1 -f -o $O/r02_knn_candidates_random \
  python bench.py --workload retrieval --bank random --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu knn random rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates_kernel -s 2 -c 1 -f -o $O/r02_knn_candidates_encoded \
  python bench.py --workload retrieval --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu knn encoded rc=$?"
ls -la $O/r02_* | tail -30
