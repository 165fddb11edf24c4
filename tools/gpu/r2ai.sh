#!/bin/bash
# GPU call r2ai (2 GPUs): the whole -m gpu suite three times in a row (flakiness check of the timing-based plausibility tests).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for k in 1 2 3; do
  timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/r2ai_pytest_$k.log 2>&1; echo "run $k rc=$? $(tail -1 $O/r2ai_pytest_$k.log)"
done
