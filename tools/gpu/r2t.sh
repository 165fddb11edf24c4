#!/bin/bash
# GPU call r2t: A-blocks per L2 band with the tile feed in place (experiment builds), configs[3] whole.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
for lib in "" libgwaspp_b200/libgwasdev_band12.so libgwaspp_b200/libgwasdev_band16.so libgwaspp_b200/libgwasdev_band24.so; do
  echo "-- configs[3] whole, lib: ${lib:-product (band 8)}" | tee -a $O/r2t_band.log
  timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --reps 2 ${lib:+--lib $lib} 2>&1 | grep "^rep" | tee -a $O/r2t_band.log
done; done
