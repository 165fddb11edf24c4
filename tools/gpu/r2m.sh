#!/bin/bash
# GPU call r2m (2 GPUs): final check -- whole -m gpu suite incl. the two-GPU tests, bench at the driver's parameters (N = 1), reference arm.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu > $O/r2m_pytest.log 2>&1; echo "rc=$?" >> $O/r2m_pytest.log; tail -6 $O/r2m_pytest.log
echo "== bench N=1, driver parameters"; /usr/bin/time -v timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench.json 2> $O/r2m_bench.err; echo "rc=$?"; grep -E "Elapsed|Maximum resident" $O/r2m_bench.err; tail -c 300 $O/r2m_bench.json
echo "== reference arm, driver parameters"; /usr/bin/time -v timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench_ref.json 2> $O/r2m_bench_ref.err; echo "rc=$?"; grep -E "Elapsed" $O/r2m_bench_ref.err; tail -c 400 $O/r2m_bench_ref.json
