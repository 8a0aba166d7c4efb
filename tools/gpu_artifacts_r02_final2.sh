#!/bin/bash
# refresh of the launch lists after the tensor-core stem + ncu --set full of the stem kernel
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_step_b32.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list-infer rc=$?"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_train_step_b16.csv python tools/profile_train_step.py 16 > gpurun_out/ncu_list_train.log 2>&1; echo "ncu-list-train rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_stem_mma_kernel -c 1 -f -o gpurun_out/r02_stem_mma_full python tools/profile_step.py 32 > gpurun_out/ncu_full_stem.log 2>&1; echo "ncu-stem rc=$?"
ls -la gpurun_out/r02_stem_mma_full.ncu-rep gpurun_out/r02_launches*.csv
