#!/bin/bash
# GPU call r2z (4 GPUs): the scaling bench's last point after the tile feed (torchrun form, as the driver launches it).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== bench N=4"; t0=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2z_bench4.json 2> $O/r2z_bench4.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 400 $O/r2z_bench4.json
