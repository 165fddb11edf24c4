#!/bin/bash
# GPU call r2aj: sanity after the last source edit -- smoke, tensor-core parity tests, one configs[3] timing.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_round2.py -q -x > $O/r2aj_pytest.log 2>&1; echo "rc=$? $(tail -1 $O/r2aj_pytest.log)"
timeout 300 python tools/time_screen.py --snps 500000 --samples 10000 --reps 2 2>&1 | grep "^rep"
