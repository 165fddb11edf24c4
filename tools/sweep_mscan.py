"""Sweeps the masked (select + scan fused) marginal kernel's configuration on configs[1]. GWASDEV_MSCAN_CFG=slots,minb,G"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libgwaspp_b200 as gw  # noqa: E402

M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
for cfg in (sys.argv[1:] or ["4,3,8", "3,3,8", "2,3,8", "4,2,8", "6,2,8", "2,4,8", "3,4,8", "2,5,8", "1,5,8", "1,6,8", "2,6,8", "2,4,16", "2,5,16"]):
    os.environ["GWASDEV_MSCAN_CFG"] = cfg
    ms = []
    for _ in range(6):
        st.select_case_control(case_mask=ca, ctrl_mask=co)      # lazy: masks only
        st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
        ms.append(st.last_scan_ms())
    t = float(np.median(ms[2:]))
    print(f"cfg {cfg:8s} {t:7.4f} ms  {M * N / 4 / t / 1e6:7.1f} GB/s")
