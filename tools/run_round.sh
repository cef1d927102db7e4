python tools/debug_r3d.py 2>/dev/null | grep -E "emb err|proj|stem|layer4.1|layer4.0.conv1.0|layer1.0.conv1.0.w" > gpurun_out/r3d_a.log
ZSV_FUSE_BN_BWD=0 ZSV_OVERLAP_WGRAD=0 python tools/debug_r3d.py 2>/dev/null | grep -E "emb err|proj|stem|layer4.1|layer4.0.conv1.0|layer1.0.conv1.0.w" > gpurun_out/r3d_b.log
echo "== default"; cat gpurun_out/r3d_a.log; echo "== no fuse, no overlap"; cat gpurun_out/r3d_b.log
