#!/bin/bash
# GPU call r2ah: G-test block size (pairs per thread block) -- 32 threads (product), 64, 128.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for lib in "" libgwaspp_b200/libgwasdev_gt64.so libgwaspp_b200/libgwasdev_gt128.so; do
  for cfg in "" "--cfg2"; do
    echo "-- lib ${lib:-product (32)} ${cfg:-cfg3}" | tee -a $O/r2ah.log
    GWASDEV_LIB=${lib:+$PWD/$lib} timeout 300 python tools/time_gtest.py $cfg 2>&1 | grep "^gtest" | tee -a $O/r2ah.log
  done
done
