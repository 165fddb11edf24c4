#!/bin/bash
# GPU call r2c (2 GPUs): binding tests, multi-device tests, torchrun bench at N = 2, band-size experiment.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/r2c_gpu.txt 2>&1
echo "== pytest binding + multi-device"; timeout 1500 python -m pytest tests/test_gpu_binding.py tests/test_gpu_round2.py -q -x -k "binding or device_table or reselection_and_row or debug_printers or multi_device or two_devices or comp_level" > $O/r2c_pytest.log 2>&1; echo "rc=$?" >> $O/r2c_pytest.log; tail -25 $O/r2c_pytest.log
echo "== torchrun bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > $O/r2c_bench_2gpu.json 2> $O/r2c_bench_2gpu.err; echo "rc=$?"; tail -c 2500 $O/r2c_bench_2gpu.json; tail -8 $O/r2c_bench_2gpu.err
echo "== band experiment (one shard of 8 at configs[3])"
for b in 8 12 16 24; do
  lib=libgwaspp_b200/libgwasdev_band$b.so; [ $b = 8 ] && lib=libgwaspp_b200/libgwasdev.so
  echo "-- band $b"; timeout 300 python tools/time_screen.py --lib $lib --snps 500000 --samples 10000 --shards 8 --reps 3 2>&1 | grep "^rep" | tee -a $O/r2c_band.log
done
echo "== band experiment configs[2]"
for b in 8 12 16 24; do
  lib=libgwaspp_b200/libgwasdev_band$b.so; [ $b = 8 ] && lib=libgwaspp_b200/libgwasdev.so
  echo "-- band $b"; timeout 300 python tools/time_screen.py --lib $lib --reps 4 2>&1 | grep "^rep" | tee -a $O/r2c_band2.log
done
