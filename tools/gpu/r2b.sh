#!/bin/bash
# GPU call r2b: new tests + binding tests, role timers of the tensor-core screen at configs[3], kernel timings, ncu captures.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
echo "== pytest round2"; timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_binding.py -q -x -s > $O/r2b_pytest.log 2>&1; echo "rc=$?" >> $O/r2b_pytest.log; tail -25 $O/r2b_pytest.log
echo "== kernels"; timeout 300 python tools/r2_kernels.py > $O/r2b_kernels.log 2>&1; grep -v trace $O/r2b_kernels.log
echo "== mma role timers, configs[3] one shard of 8 and whole"; 
GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so GWASDEV_MMA_PROF=1 timeout 600 python tools/time_screen.py --snps 500000 --samples 10000 --shards 8 --reps 3 > $O/r2b_prof_cfg3.log 2>&1; tail -12 $O/r2b_prof_cfg3.log
GWASDEV_LIB=$PWD/libgwaspp_b200/libgwasdev_sweep.so GWASDEV_MMA_PROF=1 timeout 600 python tools/time_screen.py --reps 3 > $O/r2b_prof_cfg2.log 2>&1; tail -12 $O/r2b_prof_cfg2.log
echo "== trace e2e cfg3"; timeout 600 python tools/trace_pairwise_e2e.py --cfg3 > $O/r2b_trace_cfg3.log 2>&1; tail -40 $O/r2b_trace_cfg3.log
echo "== ncu launch list (headline bench, short)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches.csv python bench.py --headline-only --no-cpu-baseline --steps 1 --warmup 3 > $O/r2b_ncu_launch.log 2>&1; echo "rc=$?"; tail -3 $O/r2b_ncu_launch.log
echo "== ncu full: one shard of the configs[3] screen"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pair_screen_mma_kernel -c 1 -o $O/r2b_prof_mma_cfg3 python tools/time_screen.py --snps 500000 --samples 10000 --shards 8 --reps 1 > $O/r2b_ncu_mma.log 2>&1; echo "rc=$?"; tail -3 $O/r2b_ncu_mma.log
echo "== ncu full: K0, K1, K1'"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_columns_kernel|marginal_scan" -c 12 -o $O/r2b_prof_marginal python tools/r2_kernels.py > $O/r2b_ncu_marginal.log 2>&1; echo "rc=$?"; tail -3 $O/r2b_ncu_marginal.log
ls -la $O | tail -20
