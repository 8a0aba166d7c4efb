#!/bin/bash
for h in "none" "det.,lane.,seg.out" "det.,lane.,seg.out,seg.d7,seg.d6"; do
  echo "HN_HILO=$h"
  HN_HILO=$h timeout 600 python bench.py --steps 20 --warmup 5 --no-latency --cpu-seconds 1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['breakdown_ms'])"
done
HN_HILO="det.,lane.,seg.out,seg.d7,seg.d6" timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -k "weight_sets or digest or full_size" --no-header -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head; cat gpurun_out/forward_parity_weight_sets.txt; cat gpurun_out/forward_640_digest_*.txt
