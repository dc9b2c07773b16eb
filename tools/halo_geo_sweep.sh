# tuning aid: RF_HALO_GEO="stacked,G,Dt,Ht,lines" forces the item shape, RF_HALO_FUSED=0/1 the MMA scheme
run() { echo "== $1 fused=$2"; RF_HALO_FUSED=$2 timeout 60 python tools/test_halo_conv.py --case $1 2>&1 | cut -c40-140,150-300; }
for c in 2048,8,32,64,56,0 2048,8,56,0,16,0 2048,8,16,0,32,0 2048,4,64,128,64,0 2048,16,8,0,16,0 8,64,16,0,16,0; do run $c 0; run $c 1; done
