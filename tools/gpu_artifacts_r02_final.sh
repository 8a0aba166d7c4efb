#!/bin/bash
# final round-2 evidence: ncu launch lists of the inference and the training step with the final kernels, ncu --set full of the
# fused grouped-conv + squeeze-excite launch, compute-sanitizer memcheck over the new kernels
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_step_b32.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list-infer rc=$?"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_train_step_b16.csv python tools/profile_train_step.py 16 > gpurun_out/ncu_list_train.log 2>&1; echo "ncu-list-train rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_se_fused_kernel -s 12 -c 2 -f -o gpurun_out/r02_gconv_se_full python tools/profile_step.py 32 > gpurun_out/ncu_full_se.log 2>&1; echo "ncu-se rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_train.py > gpurun_out/sanitize_train_memcheck.log 2>&1; echo "memcheck-train rc=$?"; grep -E "ERROR SUMMARY|sanitize_train ok" gpurun_out/sanitize_train_memcheck.log | grep -v print
bash tools/gpu_sanitize.sh 2>&1 | grep -E "rc=|ERROR SUMMARY|RACECHECK SUMMARY" | head -12
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches*.csv
