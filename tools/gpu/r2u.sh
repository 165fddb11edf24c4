#!/bin/bash
# GPU call r2u: band height from the operand row length (32 MB of A rows per band) -- parity, timings.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest mma + round2 + parity"; timeout 1500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x > $O/r2u_pytest.log 2>&1; echo "rc=$?" >> $O/r2u_pytest.log; tail -8 $O/r2u_pytest.log
echo "-- configs[3] whole" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
echo "-- configs[2]" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --reps 5 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
echo "-- configs[2] with 1 % missing calls" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --missing 0.01 --reps 4 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
echo "-- 20000/20000 samples x 20000 SNPs, 1 % missing (two-accumulator mode)" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --missing 0.01 --snps 20000 --samples 40000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
echo "-- 20000/20000 samples x 20000 SNPs, complete (split-class mode)" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --snps 20000 --samples 40000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
echo "-- 6000/6000 samples x 100000 SNPs" | tee -a $O/r2u_band.log
timeout 300 python tools/time_screen.py --snps 100000 --samples 12000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2u_band.log
