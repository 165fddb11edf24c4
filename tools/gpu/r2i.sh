#!/bin/bash
# GPU call r2i: L2 prefetch ahead of the TMA loads of the two-plane screen (experiment builds), whole configs[3] and configs[2].
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for v in 0 4 2 8 0 4; do
  lib=libgwaspp_b200/libgwasdev_pf$v.so; [ $v = 0 ] && lib=libgwaspp_b200/libgwasdev.so
  echo "-- prefetch $v stages, configs[3] whole" | tee -a $O/r2i_prefetch.log; timeout 300 python tools/time_screen.py --lib $lib --snps 500000 --samples 10000 --reps 2 2>&1 | grep "^rep" | tee -a $O/r2i_prefetch.log
done
for v in 0 4 2 8; do
  lib=libgwaspp_b200/libgwasdev_pf$v.so; [ $v = 0 ] && lib=libgwaspp_b200/libgwasdev.so
  echo "-- prefetch $v stages, configs[2]" | tee -a $O/r2i_prefetch.log; timeout 300 python tools/time_screen.py --lib $lib --reps 4 2>&1 | grep "^rep" | tee -a $O/r2i_prefetch.log
done
