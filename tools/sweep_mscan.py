"""Sweeps the configuration of the masked (select + scan fused) marginal kernel on configs[1].

Needs the sweep build of the library (extra kernel instantiations + GWASDEV_MSCAN_CFG=slots,minb,G parsing):
    python libgwaspp_b200/build.py --sweep
    GWASDEV_LIB=libgwaspp_b200/libgwasdev_sweep.so python tools/sweep_mscan.py [mode2|mode1] [cfg ...]
mode2 (default): re-selections over a table whose row totals are cached (three masked popcount streams);
mode1: the first scan of a table (three masked + three plain streams; totals caching switched off so that every scan is one)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libgwaspp_b200 as gw  # noqa: E402

args = sys.argv[1:]
mode = args.pop(0) if args and args[0] in ("mode1", "mode2") else "mode2"
M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
if mode == "mode1":
    st.set_option(gw.OPT_ROW_TOTALS, 1)
st.select_case_control(case_mask=ca, ctrl_mask=co)
st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)          # writes the row totals (mode2)
cfgs = args or ["2,4,8", "3,4,8", "4,4,8", "3,3,8", "4,3,8", "5,3,8", "6,3,8", "6,2,8", "8,2,8", "2,5,8", "4,3,16", "5,3,16"]
for cfg in cfgs:
    os.environ["GWASDEV_MSCAN_CFG"] = cfg
    ms = []
    for _ in range(8):
        st.select_case_control(case_mask=ca, ctrl_mask=co)      # lazy: masks only
        st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
        ms.append(st.last_scan_ms())
    t = float(np.median(ms[2:]))
    print(f"{mode} cfg {cfg:8s} {t:7.4f} ms  {M * N / 4 / t / 1e6:7.1f} GB/s", flush=True)
