#!/bin/bash
# round 2: ncu evidence for profiles/ (one GPU)
cd "$GRAFT_REPO_ROOT"
# 1. per-kernel launch list of one full-path step (eager launches)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu launches rc=$?"
# 2. DRAM traffic of the dominant kernel over the same command
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_conv3d_halo_kernel \
  --csv --log-file gpurun_out/r2_traffic_conv.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > gpurun_out/r2_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
# 3. --set full captures of single layers (N = 4096 patches = 16 chunks x K 4 x 64; 16 volumes for the decoder)
for c in 4096,8,32,64,56,0 4096,16,8,0,16,0 4096,4,64,128,64,0 16,64,16,0,16,0; do
  n=$(echo $c | tr ',' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3d_halo_kernel -c 1 -f -o gpurun_out/r2_halo_$n \
    python tools/test_halo_conv.py --case $c > gpurun_out/r2_halo_$n.log 2>&1; echo "ncu full $c rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_norm_split_halo_kernel -c 1 -f -o gpurun_out/r2_split_4096_8_32_64 \
  python tools/test_halo_conv.py --case 4096,8,32,64,56,0 > /dev/null 2>&1; echo "ncu split rc=$?"
ls -la gpurun_out/*.ncu-rep | tail
