#!/bin/bash
# GPU call r2d (ncu only): whole configs[3] screen on one GPU (roofline.traffic of the headline), the masked scans, the
# four-plane screen at configs[2] with missing calls.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
echo "== ncu full: whole configs[3] screen (one launch, ~40 replays of 3.3 s)"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma_kernel -c 1 -o $O/r2d_prof_mma_cfg3_whole python tools/time_screen.py --snps 500000 --samples 10000 --reps 1 > $O/r2d_ncu_mma.log 2>&1; echo "rc=$?"; tail -3 $O/r2d_ncu_mma.log
echo "== ncu full: masked scans (MODE 1 first scan, MODE 2 re-selections)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:marginal_scan_masked_kernel -c 4 -o $O/r2d_prof_masked python tools/sweep_mscan.py mode2 5,3,16 > $O/r2d_ncu_masked.log 2>&1; echo "rc=$?"; tail -3 $O/r2d_ncu_masked.log
echo "== ncu full: four-plane screen, configs[2] with 1 % missing calls"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma4_kernel -c 1 -o $O/r2d_prof_mma4 python tools/time_screen.py --missing 0.01 --reps 1 > $O/r2d_ncu_mma4.log 2>&1; echo "rc=$?"; tail -3 $O/r2d_ncu_mma4.log
ls -la $O/r2d_*
