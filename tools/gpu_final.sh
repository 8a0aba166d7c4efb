#!/bin/bash
# what the driver runs at round end, plus the artefacts kept under profiles/
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --dump-ops 2>&1 | tail -1 > gpurun_out/bench_default.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read()); print('infer', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_b1_ms'], 'roof', d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-200
timeout 600 python bench.py --mode train 2>&1 | tail -1 | tee gpurun_out/bench_train_default.json | cut -c1-200
timeout 600 python bench.py --mode train --batch 32 --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_train_b32.json | cut -c1-200
