#!/bin/bash
# GPU call r2w (ncu only, after the same commands ran without ncu in r2r-r2v): final state of the tensor-core screens.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
echo "== ncu metrics: whole configs[3] pass"
timeout 900 ncu --metrics $M --clock-control none -k regex:pair_screen_mma_kernel -c 1 --csv --log-file $O/r2w_ncu_cfg3_whole.csv python tools/time_screen.py --snps 500000 --samples 10000 --reps 1 > $O/r2w_ncu1.log 2>&1; echo "rc=$?"; tail -8 $O/r2w_ncu_cfg3_whole.csv | cut -d, -f13-15
echo "== ncu full: one shard of 8 of the configs[3] screen"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma_kernel -c 1 -o $O/r2w_prof_mma_cfg3 python tools/time_screen.py --snps 500000 --samples 10000 --shards 8 --reps 1 > $O/r2w_ncu2.log 2>&1; echo "rc=$?"; tail -2 $O/r2w_ncu2.log
echo "== ncu full: configs[2], two-plane and (1 % missing) four-plane kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma_kernel -c 1 -o $O/r2w_prof_mma_cfg2 python tools/time_screen.py --reps 1 > $O/r2w_ncu3.log 2>&1; echo "rc=$?"; tail -2 $O/r2w_ncu3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma4_kernel -c 1 -o $O/r2w_prof_mma4 python tools/time_screen.py --missing 0.01 --reps 1 > $O/r2w_ncu4.log 2>&1; echo "rc=$?"; tail -2 $O/r2w_ncu4.log
echo "== ncu launch list (headline bench, short)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2w_launches.csv python bench.py --headline-only --no-cpu-baseline --steps 1 --warmup 3 > $O/r2w_ncu_launch.log 2>&1; echo "rc=$?"; tail -2 $O/r2w_ncu_launch.log | cut -c1-300
ls -la $O | grep r2w
