T1="22,16,56,56,144,64,3,1,1,1,1,1,1,0,0"
echo "== aligned box 8,1,1,8 (rows 64)"; ZSV_DEBUG_HALO_BOX=8,1,1,8 ZSV_DEBUG_EPI=7 python tools/bench_conv.py $T1 2>&1 | cut -c1-150
echo "== unaligned box 4,1,1,16 (rows 64)"; ZSV_DEBUG_HALO_UNALIGNED=1 ZSV_DEBUG_HALO_BOX=4,1,1,16 ZSV_DEBUG_EPI=7 python tools/bench_conv.py $T1 2>&1 | cut -c1-150
echo "== unaligned box 2,2,1,16 (rows 64)"; ZSV_DEBUG_HALO_UNALIGNED=1 ZSV_DEBUG_HALO_BOX=2,2,1,16 ZSV_DEBUG_EPI=7 python tools/bench_conv.py $T1 2>&1 | cut -c1-150
echo "== aligned box 4,2,1,8 (rows 64)"; ZSV_DEBUG_HALO_BOX=4,2,1,8 ZSV_DEBUG_EPI=7 python tools/bench_conv.py $T1 2>&1 | cut -c1-150
