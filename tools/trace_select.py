import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import libgwaspp_b200 as gw
M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    torch.cuda.synchronize(); print("call ms", (time.perf_counter() - t0) * 1e3)
L = st.L
import ctypes as C
for _ in range(4):
    t0 = time.perf_counter()
    L.gwasdev_select_case_control(st.h, gw._ptr(ca), gw._ptr(co))
    print("raw C call ms", (time.perf_counter() - t0) * 1e3)
