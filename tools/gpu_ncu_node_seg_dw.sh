#!/bin/bash
# ncu --set full captures: BiFPN P3 node kernel and the narrow-N seg convs (d6, d7.p00..p11, out)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_node_kernel -s 3 -c 1 -f -o gpurun_out/prof_node4 python tools/profile_step.py 32 > gpurun_out/ncu_full_node.log 2>&1; echo "ncu-node rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_conv_gemm_kernel -s 198 -c 6 -f -o gpurun_out/prof_seg_narrow python tools/profile_step.py 32 > gpurun_out/ncu_full_seg.log 2>&1; echo "ncu-seg rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_dw_multi_kernel -s 0 -c 1 -f -o gpurun_out/prof_dwm python tools/profile_step.py 32 > gpurun_out/ncu_full_dwm.log 2>&1; echo "ncu-dwm rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
