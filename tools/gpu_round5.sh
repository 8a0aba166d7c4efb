#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_postproc.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-160
timeout 600 python bench.py --steps 10 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_infer.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_infer.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['latency_b1_ms']['p50'], d['roofline']['frac'], d['roofline']['traffic'])"
