#!/bin/bash
# GPU call r2ad (1 GPU): final regression -- smoke, whole -m gpu suite, bench at the driver's parameters, reference arm.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu > $O/r2ad_pytest.log 2>&1; echo "rc=$?" >> $O/r2ad_pytest.log; tail -5 $O/r2ad_pytest.log
echo "== bench N=1, driver parameters"; t0=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2ad_bench.json 2> $O/r2ad_bench.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 200 $O/r2ad_bench.json
echo "== reference arm"; t0=$(date +%s); timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2ad_bench_ref.json 2> $O/r2ad_bench_ref.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 200 $O/r2ad_bench_ref.json
