#!/bin/bash
# GPU call r2n (1 GPU): bench at the driver's parameters (N = 1) and the reference arm, wall time of each.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1, driver parameters"; t0=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2n_bench.json 2> $O/r2n_bench.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 300 $O/r2n_bench.json
echo "== reference arm, driver parameters"; t0=$(date +%s); timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2n_bench_ref.json 2> $O/r2n_bench_ref.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 400 $O/r2n_bench_ref.json
