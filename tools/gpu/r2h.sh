#!/bin/bash
# GPU call r2h (8 GPUs): multi-device tests, torchrun bench at N = 8 and N = 4 (the driver's SCALE path).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=${1:-8}
echo "== pytest multi-device"; timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "multi_device or two_devices" > $O/r2h_pytest.log 2>&1; echo "rc=$?" >> $O/r2h_pytest.log; tail -5 $O/r2h_pytest.log
for n in $N 4; do
  echo "== torchrun bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > $O/r2h_bench_${n}gpu.json 2> $O/r2h_bench_${n}gpu.err; echo "rc=$?"; tail -c 1200 $O/r2h_bench_${n}gpu.json; tail -4 $O/r2h_bench_${n}gpu.err
done
