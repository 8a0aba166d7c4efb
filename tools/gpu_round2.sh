#!/bin/bash
tools/gpu_ops_tests.sh tests/test_gpu_train_ops.py 200 | tr '\t' ' ' | grep -E "passed|failed" | tr '\n' ' '; echo
tools/gpu_ops_tests.sh tests/test_gpu_train.py 900 | tr '\t' ' ' | grep -E "passed|failed" | tr '\n' ' '; echo
timeout 600 python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -3 > gpurun_out/bench_train.log; cut -c1-200 gpurun_out/bench_train.log
HN_SIDE_WGRAD=0 timeout 600 python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
