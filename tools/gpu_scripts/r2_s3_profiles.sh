#!/bin/bash
# round 2, session 3: tests + bench lines + ncu evidence for profiles/ (one GPU)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > $O/r02s3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02s3_pytest_gpu.log
timeout 900 python bench.py > $O/r02s3_bench_full_n1.json 2> $O/r02s3_bench_full_n1.err; echo "full rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02s3_bench_reference_arm.json 2> $O/r02s3_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py --workload retrieval --no-cpu-baseline > $O/r02s3_bench_retrieval_encoded.json 2>/dev/null; echo "retr rc=$?"
timeout 600 python bench.py --workload surface --no-cpu-baseline > $O/r02s3_bench_surface_n1.json 2>/dev/null; echo "surface rc=$?"
timeout 900 python bench.py --workload stages > $O/r02s3_stage_rooflines.json 2>/dev/null; echo "stages rc=$?"
# launch list + DRAM traffic of the dominant kernel over one eager full-path step
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02s3_launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/print_last_step.py $O/r02s3_launches_full.csv > $O/r02s3_full_path_last_step.txt; tail -1 $O/r02s3_full_path_last_step.txt
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_conv3d_halo_kernel \
  --csv --log-file $O/r02s3_traffic_conv.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu traffic rc=$?"
# --set full captures: the 96 -> 56 join, the W-pair 8 -> 16 @ 16^3 layer with the pooling epilogue (inside the fused-front DoubleConv), the fused front kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3d_halo_kernel -c 1 -f -o $O/r02s3_halo_4096_8_32_64_56_0 \
  python tools/test_halo_conv.py --case 4096,8,32,64,56,0 > /dev/null 2>&1; echo "ncu full join rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3d_halo_kernel -s 1 -c 1 -f -o $O/r02s3_halo_wp_8to16_16cube \
  python tools/front_time.py 4096 > /dev/null 2>&1; echo "ncu full wp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unet_front16_kernel -s 1 -c 1 -f -o $O/r02s3_unet_front16 \
  python tools/front_time.py 4096 > /dev/null 2>&1; echo "ncu full front rc=$?"
ls -la $O/r02s3_* | tail -30
