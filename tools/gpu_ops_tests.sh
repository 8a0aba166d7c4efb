#!/bin/bash
# Run every test function of a GPU test file in its own process (a faulting kernel cannot poison the next test).
# usage: tools/gpu_ops_tests.sh tests/test_gpu_train_ops.py [per-test timeout seconds]
f=$1; to=${2:-180}
mkdir -p gpurun_out
log=gpurun_out/$(basename $f .py).log
: > $log
for t in $(grep -o "^def test_[a-z0-9_]*" $f | sed 's/def //'); do
  echo "=== $t" >> $log
  timeout $to python -m pytest "$f::$t" -q -x --no-header -p no:cacheprovider 2>&1 | tail -25 >> $log
  echo "exit $?" >> $log
done
grep -E "^===|passed|failed|error|exit" $log | paste - - - | head -60
