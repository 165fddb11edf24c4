#!/bin/bash
# GPU call r2s: tile feed in both tensor-core kernels -- whole GPU suite, timings, DRAM traffic of a whole configs[3] pass.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu -x > $O/r2s_pytest.log 2>&1; echo "rc=$?" >> $O/r2s_pytest.log; tail -6 $O/r2s_pytest.log
for miss in 0.01 0.0000001; do
  echo "-- configs[2] with missing rate $miss" | tee -a $O/r2s_feed.log
  timeout 300 python tools/time_screen.py --missing $miss --reps 4 2>&1 | grep "^rep" | tee -a $O/r2s_feed.log
done
echo "-- 20000/20000 samples x 20000 SNPs, 1 % missing (two-accumulator mode)" | tee -a $O/r2s_feed.log
timeout 300 python tools/time_screen.py --missing 0.01 --snps 20000 --samples 40000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2s_feed.log
echo "-- 20000/20000 samples x 20000 SNPs, complete (split-class mode)" | tee -a $O/r2s_feed.log
timeout 300 python tools/time_screen.py --snps 20000 --samples 40000 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2s_feed.log
echo "== ncu: DRAM traffic of a whole configs[3] pass"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k regex:pair_screen_mma_kernel -c 1 --csv --log-file $O/r2s_ncu_cfg3.csv python tools/time_screen.py --snps 500000 --samples 10000 --reps 1 > $O/r2s_ncu.log 2>&1; echo "rc=$?"
cat $O/r2s_ncu_cfg3.csv | tail -8
