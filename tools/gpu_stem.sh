#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu --no-header -p no:cacheprovider -k "stem or lockstep or golden or digest" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_train_ops.py -x -q -m gpu --no-header -p no:cacheprovider -k "stem" 2>&1 | tail -2
for v in 0 1; do
  HN_STEM_MMA=$v timeout 600 python bench.py --dump-ops 2>&1 | tail -1 > gpurun_out/bench_stem$v.json
  python -c "
import json; d=json.loads(open('gpurun_out/bench_stem$v.json').read()); print('HN_STEM_MMA=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_b1_ms']['p50'], d['roofline']['hbm_bound_kernels']['stem'])"
done
