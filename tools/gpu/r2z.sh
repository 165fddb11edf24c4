#!/bin/bash
# GPU call r2z (8 GPUs): the scaling bench's last point after the tile feed (torchrun form, as the driver launches it).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== bench N=8"; t0=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2z_bench8.json 2> $O/r2z_bench8.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 400 $O/r2z_bench8.json
