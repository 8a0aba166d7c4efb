#!/bin/bash
# the fused squeeze-excite launches: unit tests, lockstep, then A/B of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu --no-header -p no:cacheprovider -k "se_fused or gconv_se or lockstep or determinism or graph_replay or golden" 2>&1 | tail -8
for v in 0 1; do
  HN_GCONV_SE=$v timeout 600 python bench.py --dump-ops 2>&1 | tail -1 > gpurun_out/bench_gconvse$v.json
  cp gpurun_out/op_times.txt gpurun_out/op_times_gconvse$v.txt
  python -c "
import json; d=json.loads(open('gpurun_out/bench_gconvse$v.json').read()); print('HN_GCONV_SE=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_b1_ms']['p50'], 'launches', d.get('gpu_launches'), 'roof', d['roofline']['frac'])"
done
grep -E "c2se|s2.b1.se" gpurun_out/op_times_gconvse1.txt | head -30
