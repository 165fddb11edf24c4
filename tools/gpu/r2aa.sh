#!/bin/bash
# GPU call r2aa: the tile-feed coverage test, then the whole -m gpu suite once more.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py -q -x -k tile_feed > $O/r2aa_feed.log 2>&1; echo "rc=$?" >> $O/r2aa_feed.log; tail -15 $O/r2aa_feed.log
timeout 2400 python -m pytest tests -q -m gpu > $O/r2aa_pytest.log 2>&1; echo "rc=$?" >> $O/r2aa_pytest.log; tail -5 $O/r2aa_pytest.log
