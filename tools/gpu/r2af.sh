#!/bin/bash
# GPU call r2af: A operand of the missing-call kernel through a box whose fourth row per SNP is out of bounds (zero-filled, not read).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "-- configs[2] with 1 % missing calls" | tee -a $O/r2af.log
timeout 120 python tools/time_screen.py --missing 0.01 --reps 4 2>&1 | grep -E "^rep|rror|trap" | tee -a $O/r2af.log
echo "== pytest mma + round2 + parity"; timeout 1500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x > $O/r2af_pytest.log 2>&1; echo "rc=$?" >> $O/r2af_pytest.log; tail -6 $O/r2af_pytest.log
echo "-- 20000/20000 samples x 20000 SNPs, 1 % missing (two-accumulator mode)" | tee -a $O/r2af.log
timeout 300 python tools/time_screen.py --missing 0.01 --snps 20000 --samples 40000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2af.log
