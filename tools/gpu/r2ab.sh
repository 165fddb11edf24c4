#!/bin/bash
# GPU call r2ab: role timers of the two-plane kernel (sweep build) with the tile feed, configs[3] whole and configs[2].
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so GWASDEV_MMA_PROF=1 timeout 600 python tools/time_screen.py --snps 500000 --samples 10000 --reps 2 > $O/r2ab_prof_cfg3.log 2>&1; grep -E "prof|^rep" $O/r2ab_prof_cfg3.log
GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so GWASDEV_MMA_PROF=1 timeout 600 python tools/time_screen.py --reps 3 > $O/r2ab_prof_cfg2.log 2>&1; grep -E "prof|^rep" $O/r2ab_prof_cfg2.log
