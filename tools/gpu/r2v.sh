#!/bin/bash
# GPU call r2v (2 GPUs): whole -m gpu suite after the tile feed / band changes, then the 2-GPU bench (torchrun form).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu > $O/r2v_pytest.log 2>&1; echo "rc=$?" >> $O/r2v_pytest.log; tail -6 $O/r2v_pytest.log
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2v_bench2.json 2> $O/r2v_bench2.err; echo "rc=$?"; tail -c 600 $O/r2v_bench2.json | head -c 600
