#!/bin/bash
# round-2 artefacts for profiles/: ncu launch lists (time + DRAM bytes) of one inference step and one training step, ncu --set full
# of the weight-gradient kernel (seg.d1: 624->512 3x3 @40^2, and a stage-4 1x1) and of the BatchNorm backward kernels.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_step_b32.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list-infer rc=$?"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_train_step_b16.csv python tools/profile_train_step.py 16 > gpurun_out/ncu_list_train.log 2>&1; echo "ncu-list-train rc=$?"
HN_NVTX=1 HN_SIDE_WGRAD=0 timeout 900 ncu --profile-from-start off --nvtx --nvtx-include "wgrad:seg.d1/" --nvtx-include "wgrad:backbone.s4.b5.c1/" --set full --clock-control none --import-source on -k regex:hn_conv_wgrad_kernel -f -o gpurun_out/r02_wgrad_seg_d1_s4_c1_full python tools/profile_train_step.py 16 > gpurun_out/ncu_full_wgrad.log 2>&1; echo "ncu-wgrad rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"hn_bn_bwd_apply_kernel|hn_red1_kernel" -s 40 -c 4 -f -o gpurun_out/r02_bn_bwd_full python tools/profile_train_step.py 16 > gpurun_out/ncu_full_bn.log 2>&1; echo "ncu-bn rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
