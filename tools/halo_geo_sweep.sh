run() { echo "== $1 geo=$2"; RF_HALO_GEO=$2 timeout 60 python tools/test_halo_conv.py --case $1 2>&1 | cut -c40-60,150-300; }
for g in 1,1,8,8,1 0,1,4,8,0 0,1,4,8,1; do run 2048,8,32,64,56,0 $g; done
for g in 0,1,8,16,1 0,1,4,16,1 0,1,2,16,1 0,1,1,16,1 0,1,2,16,0; do run 2048,16,8,0,16,0 $g; done
for g in 0,1,2,8,0 0,1,1,8,0 0,1,4,4,0; do run 8,64,16,0,16,0 $g; done
for g in 1,1,8,8,1 0,1,4,8,1; do run 2048,8,16,0,32,0 $g; done
for g in 1,2,4,4,0 1,1,4,4,0; do run 2048,4,64,128,64,0 $g; done
