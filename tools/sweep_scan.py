"""Tuning sweep of the marginal-scan kernel configuration (GWASDEV_SCAN_CFG=slots,blocks_per_sm) on configs[1]."""
import os, sys, subprocess, json
cfgs = sys.argv[1:] or ["5,3,8,0", "5,3,8,1", "5,3,16,0", "5,3,16,1", "5,3,32,0", "5,3,32,1", "3,4,32,1", "2,5,32,1", "2,5,8,1"]
for c in cfgs:
    env = dict(os.environ, GWASDEV_SCAN_CFG=c)
    out = subprocess.run([sys.executable, "bench.py", "--no-pairwise", "--no-cpu-baseline", "--steps", "20"], env=env,
                         capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(c, "FAILED", out.stderr[-300:]); continue
    d = json.loads(line[-1])
    print(c, d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], flush=True)
