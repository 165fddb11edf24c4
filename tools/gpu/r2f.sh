#!/bin/bash
# GPU call r2f/r2g: compute-sanitizer on smoke(), ONE tool per call (B200_PROFILING.md). usage: r2f.sh memcheck|racecheck
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
T=$1
timeout 1500 compute-sanitizer --tool $T --print-limit 30 --log-file $O/r2_sanitizer_$T.log python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_${T}_stdout.log 2>&1
echo "rc=$?" >> $O/r2_sanitizer_${T}_stdout.log
tail -5 $O/r2_sanitizer_${T}_stdout.log; tail -15 $O/r2_sanitizer_$T.log
nvidia-smi --query-gpu=name,memory.used --format=csv
