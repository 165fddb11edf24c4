"""A/B timing of two builds of libgwasdev.so on the configs[1] marginal scan: python tools/ab_marginal.py LIB [LIB ...]
(each library in its own subprocess; prints the scan kernel's CUDA-event time, compacted and masked)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import libgwaspp_b200 as gw
gw.LIB_PATH = os.path.abspath(sys.argv[1])
M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N); st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE); ca, co = gw.stream_masks(pheno)
dc = torch.empty((M, 8), dtype=torch.int32, device="cuda"); ds = torch.empty((M, 8), dtype=torch.float64, device="cuda")
masked = []
for _ in range(6):
    st.select_case_control(case_mask=ca, ctrl_mask=co); st.marginal_scan_into(0, M, counts=dc, stats=ds); masked.append(st.last_scan_ms())
st.marginal_scan_into(0, M, counts=dc, stats=ds)
comp = []
for _ in range(30):
    st.marginal_scan_into(0, M, counts=dc, stats=ds); comp.append(st.last_scan_ms())
print(f"{sys.argv[1]}: compacted {np.median(comp):.4f} ms (min {min(comp):.4f})  masked {np.median(masked[2:]):.4f} ms")
''' % ROOT
for lib in sys.argv[1:]:
    for rep in range(2):
        subprocess.run([sys.executable, "-c", CHILD, lib], check=False)
