#!/bin/bash
# GPU call r2l: sparse plane choice -- parity tests of the tensor-core engine, then A/B against the classic planes.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest mma + round2"; timeout 1500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x > $O/r2l_pytest.log 2>&1; echo "rc=$?" >> $O/r2l_pytest.log; tail -8 $O/r2l_pytest.log
for rep in 1 2; do
for mode in "" "--classic-planes"; do
  echo "-- configs[3] whole, planes: ${mode:-rarest two}" | tee -a $O/r2l_planes.log
  timeout 300 python tools/time_screen.py $mode --snps 500000 --samples 10000 --reps 2 2>&1 | grep "^rep" | tee -a $O/r2l_planes.log
done; done
for mode in "" "--classic-planes"; do
  echo "-- configs[2], planes: ${mode:-rarest two}" | tee -a $O/r2l_planes.log
  timeout 300 python tools/time_screen.py $mode --reps 4 2>&1 | grep "^rep" | tee -a $O/r2l_planes.log
done
