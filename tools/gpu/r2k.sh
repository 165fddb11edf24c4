#!/bin/bash
# GPU call r2k: the whole -m gpu suite, smoke, and a short default bench after the raw-row operand change.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "rc=$?" >> $O/r2k_smoke.log; tail -2 $O/r2k_smoke.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu -x > $O/r2k_pytest.log 2>&1; echo "rc=$?" >> $O/r2k_pytest.log; tail -15 $O/r2k_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2k_bench.json 2> $O/r2k_bench.err; echo "rc=$?"; tail -c 600 $O/r2k_bench.json; tail -5 $O/r2k_bench.err
