#!/bin/bash
# GPU call r2q: SM clock and board power while the missing-call tensor-core kernel runs back to back.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active,temperature.gpu --format=csv,noheader -lms 100 > $O/r2q_smi.log &
SMI=$!
sleep 1
timeout 300 python tools/time_screen.py --missing 0.01 --reps 60 2>&1 | grep "^rep" > $O/r2q_reps.log
sleep 1
timeout 300 python tools/time_screen.py --reps 150 2>&1 | grep "^rep" > $O/r2q_reps_clean.log
kill $SMI
head -3 $O/r2q_reps.log; tail -3 $O/r2q_reps.log; tail -2 $O/r2q_reps_clean.log
awk -F, '{print $1, $3, $4}' $O/r2q_smi.log | sort | uniq -c | sort -k1 -n -r | head -40
