timeout 900 python -m pytest tests/test_gpu_train_ops.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -1
for c in 296 592 1184; do HN_RED_CHUNKS=$c timeout 600 python bench.py --mode train 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('chunks $c', d['ms_per_step'], d['value'])"; done
