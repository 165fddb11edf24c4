#!/bin/bash
# GPU call r2r: tile feed (CTA pairs draw tiles from a device counter) -- parity, then timing at configs[3] / configs[2].
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest mma + round2 + parity"; timeout 1500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x > $O/r2r_pytest.log 2>&1; echo "rc=$?" >> $O/r2r_pytest.log; tail -8 $O/r2r_pytest.log
echo "-- configs[3] whole" | tee -a $O/r2r_feed.log
timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2r_feed.log
echo "-- configs[3], shard 0 of 8" | tee -a $O/r2r_feed.log
timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --shards 8 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2r_feed.log
echo "-- configs[2]" | tee -a $O/r2r_feed.log
timeout 300 python tools/time_screen.py --reps 5 2>&1 | grep "^rep" | tee -a $O/r2r_feed.log
