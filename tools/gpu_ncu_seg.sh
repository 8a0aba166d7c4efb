#!/bin/bash
# ncu --set full of seg decoder.2 and decoder.3 (parity 00): the two launches after index conv_index[seg.d2]
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py 32 > /dev/null 2>&1   # writes gpurun_out/conv_index.txt
idx=$(awk '$1=="seg.d2"{print $2}' gpurun_out/conv_index.txt)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_conv_gemm_kernel -s $idx -c 2 -f -o gpurun_out/prof_seg_d2_d3 python tools/profile_step.py 32 > gpurun_out/ncu_full_seg.log 2>&1; echo "ncu-seg idx=$idx rc=$?"
