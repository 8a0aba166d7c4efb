#!/bin/bash
# One GPU session of the development loop: training tests, forward / demo tests, train bench + kernel profile, inference bench.
tools/gpu_ops_tests.sh tests/test_gpu_train.py 900
timeout 600 python -m pytest tests/test_gpu_demo.py tests/test_gpu_forward.py -q --no-header -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -3 > gpurun_out/bench_train.log; cat gpurun_out/bench_train.log
timeout 600 python tools/profile_train.py > /dev/null 2>gpurun_out/prof_err.log; head -100 gpurun_out/train_profile.txt | tail -60
timeout 900 python bench.py --steps 10 --warmup 5 --dump-ops 2>&1 | tail -2 > gpurun_out/bench_infer.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_infer.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['latency_b1_ms'], d['roofline']['frac'])"
