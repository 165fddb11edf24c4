#!/bin/bash
# GPU call r2y: band height sweep with the tile feed, finer (experiment builds with a forced band), configs[3] whole.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
for b in 10 12 13 14 15 16; do
  echo "-- configs[3] whole, band $b" | tee -a $O/r2y_band.log
  timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --reps 2 --lib libgwaspp_b200/libgwasdev_band$b.so 2>&1 | grep "^rep 1" | tee -a $O/r2y_band.log
done; done
