#!/bin/bash
# GPU call r2a: tests, short bench, kernel timings, K1' sweep. Every stage under its own timeout; later stages run even if
# an earlier one fails. Logs under gpurun_out/r2a_*.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
echo "== i8 peak probe"; timeout 120 python -c "
import libgwaspp_b200 as gw
print('i8 peak (burst, sustained) TOP/s:', gw.i8_peak(0))
print('popc peak:', gw.popc_peak(0))
" > $O/r2a_peak.log 2>&1; echo "rc=$?" >> $O/r2a_peak.log; tail -3 $O/r2a_peak.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; echo "rc=$?" >> $O/r2a_smoke.log; tail -2 $O/r2a_smoke.log
echo "== pytest new"; timeout 1500 python -m pytest tests/test_gpu_round2.py -q -x -s > $O/r2a_pytest_new.log 2>&1; echo "rc=$?" >> $O/r2a_pytest_new.log; tail -15 $O/r2a_pytest_new.log
echo "== pytest old"; timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_round2.py > $O/r2a_pytest_old.log 2>&1; echo "rc=$?" >> $O/r2a_pytest_old.log; tail -15 $O/r2a_pytest_old.log
echo "== kernels"; timeout 300 python tools/r2_kernels.py > $O/r2a_kernels.log 2>&1; echo "rc=$?" >> $O/r2a_kernels.log; cat $O/r2a_kernels.log
echo "== sweep"; GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so timeout 300 python tools/sweep_mscan.py mode2 > $O/r2a_sweep2.log 2>&1; cat $O/r2a_sweep2.log
GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so timeout 300 python tools/sweep_mscan.py mode1 2,4,8 3,4,8 3,3,8 4,3,8 > $O/r2a_sweep1.log 2>&1; cat $O/r2a_sweep1.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "rc=$?"; tail -c 1500 $O/r2a_bench.json; tail -5 $O/r2a_bench.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --headline-only > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err; echo "rc=$?"; tail -c 600 $O/r2a_bench_ref.json
