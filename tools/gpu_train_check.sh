#!/bin/bash
# training kernels: operator tests, whole-step tests, then the training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --mode train 2>&1 | tail -1 | tee gpurun_out/bench_train_default.json | cut -c1-220
