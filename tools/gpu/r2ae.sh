#!/bin/bash
# GPU call r2ae: the whole default bench once more after the last bench.py edits (short).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
t0=$(date +%s); timeout 1200 python bench.py --steps 3 --warmup 3 > $O/r2ae_bench.json 2> $O/r2ae_bench.err; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -3 $O/r2ae_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ae_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], json.dumps(d['roofline']['l2_to_sm'])[:200])
print(json.dumps(d['pairwise_missing_calls']['tensor_cores_four_planes'])[:600])
PY
