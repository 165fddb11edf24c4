#!/bin/bash
# GPU call r2x (1 GPU): smoke + bench at the driver's parameters after the tile feed / band / missing-call kernel changes.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1, driver parameters"; t0=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 300 $O/r2x_bench.json
