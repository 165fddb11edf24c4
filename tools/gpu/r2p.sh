#!/bin/bash
# GPU call r2p: where the missing-call tensor-core kernel spends its time (experiment builds: no statistic / no epilogue at all).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for lib in "" libgwaspp_b200/libgwasdev_m4x1.so libgwaspp_b200/libgwasdev_m4x2.so; do
  echo "-- configs[2] with 1 % missing calls, lib: ${lib:-product}" | tee -a $O/r2p_m4.log
  timeout 300 python tools/time_screen.py --missing 0.01 --reps 4 ${lib:+--lib $lib} 2>&1 | grep "^rep" | tee -a $O/r2p_m4.log
done
