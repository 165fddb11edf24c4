#!/bin/bash
# GPU call r2e: band height of the tile schedule on the WHOLE configs[3] problem (one GPU), plus K1' check after the lane fix.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for b in 8 16 32 12 24; do
  lib=libgwaspp_b200/libgwasdev_band$b.so; [ $b = 8 ] && lib=libgwaspp_b200/libgwasdev.so
  echo "-- band $b" | tee -a $O/r2e_band_whole.log; timeout 300 python tools/time_screen.py --lib $lib --snps 500000 --samples 10000 --reps 2 2>&1 | grep "^rep" | tee -a $O/r2e_band_whole.log
done
timeout 300 python tools/r2_kernels.py 2>&1 | grep -v trace | tee $O/r2e_kernels.log
