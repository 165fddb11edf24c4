"""Workload for the round-1m ncu captures: device-side TPED load (configs[0] size), the fused select+scan on the raw
rows, K0 and the compacted scan at configs[1]. Prints kernel times (CUDA events inside the library)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libgwaspp_b200 as gw  # noqa: E402
from test_gpu_ingest import tped_bytes  # noqa: E402

# ---- ingest: 10 000 SNPs x 2 000 samples of TPED text
M0, N0 = 10_000, 2_000
codes = np.random.default_rng(1).choice(4, size=(M0, N0), p=[0.62, 0.3, 0.07, 0.01]).astype(np.uint8)
path = os.path.join(tempfile.mkdtemp(), "c.tped")
open(path, "wb").write(tped_bytes(codes))
with gw.GenoStore(M0, N0) as st:
    for _ in range(3):
        assert st.load_tped(path) == M0

# ---- scans at configs[1]
M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
for rep in range(3):
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)          # masked scan on the raw rows
    t_masked = st.last_scan_ms()
    st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)          # K0, then the compacted scan
    st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
    t_both = st.last_scan_ms()
    st.marginal_scan_into(0, M, counts=d_counts)                         # counts only: what the fp64 epilogue costs
    t_counts = st.last_scan_ms()
    print(f"rep {rep}: masked scan {t_masked:.4f} ms | compacted scan counts+stats {t_both:.4f} ms | counts only {t_counts:.4f} ms")
print("hbm read probe GB/s:", gw.hbm_read_peak(0, 1 << 30))
