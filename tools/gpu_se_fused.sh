#!/bin/bash
# the fused squeeze-excite launch: unit test, lockstep, postproc (shorter cell sort), then A/B of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu --no-header -p no:cacheprovider -k "se_fused or lockstep or determinism or graph_replay" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_postproc.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
for v in 0 1; do
  HN_SE_FUSED=$v timeout 600 python bench.py --dump-ops 2>&1 | tail -1 > gpurun_out/bench_sefused$v.json
  cp gpurun_out/op_times.txt gpurun_out/op_times_sefused$v.txt
  python -c "
import json; d=json.loads(open('gpurun_out/bench_sefused$v.json').read()); print('HN_SE_FUSED=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_b1_ms']['p50'], 'launches', d.get('gpu_launches'))"
done
grep -E "se_fused|s4.b3.se|s3.b3.se|s2.b1.se" gpurun_out/op_times_sefused1.txt | head -12
