#!/bin/bash
# GPU call r2ag: TMA + MMA pipeline of the missing-call kernel alone (epilogue compiled out) with the zero-filled A padding rows.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for lib in libgwaspp_b200/libgwasdev_m4x2.so ""; do
echo "-- configs[2] with 1 % missing calls, lib ${lib:-product}" | tee -a $O/r2ag.log
timeout 120 python tools/time_screen.py --missing 0.01 --reps 4 ${lib:+--lib $lib} 2>&1 | grep -E "^rep" | tee -a $O/r2ag.log
done
