#!/bin/bash
# GPU call r2ac: L2 -> SM read rate probe at several buffer sizes, then a short headline bench carrying roofline.l2_to_sm.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python - <<'PY' 2>&1 | tee $O/r2ac_l2.log
import libgwaspp_b200 as gw
for mb in (4, 8, 16, 32, 48, 64):
    for passes in (100, 12800 // mb):
        print(f"L2-resident buffer {mb:3d} MiB, {passes:5d} passes: {gw.l2_read_peak(0, mb << 20, passes) / 1e3:6.2f} TB/s")
print(f"HBM read (2 GiB): {gw.hbm_read_peak(0) / 1e3:.2f} TB/s")
PY
timeout 900 python bench.py --headline-only --no-cpu-baseline --steps 3 --warmup 3 > $O/r2ac_bench.json 2> $O/r2ac_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], json.dumps(d['roofline']['l2_to_sm']))
PY
