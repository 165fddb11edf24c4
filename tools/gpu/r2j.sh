#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 python tools/ab_k1.py libgwaspp_b200/libgwasdev_r1.so libgwaspp_b200/libgwasdev.so 2>&1 | tee $O/r2j_ab_k1.log
