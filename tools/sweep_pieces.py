import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import libgwaspp_b200 as gw
M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N); st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE); ca, co = gw.stream_masks(pheno)
hc = torch.empty((M, 8), dtype=torch.int32, pin_memory=True); hs = torch.empty((M, 8), dtype=torch.float64, pin_memory=True)
for pieces in (1, 2, 3, 4, 6, 8):
    st.set_option(gw.OPT_SCAN_PIECES, pieces)
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st.select_case_control(case_mask=ca, ctrl_mask=co)
        st.marginal_scan_into(0, M, counts=hc, stats=hs, on_device=False)
        ts.append(time.perf_counter() - t0)
    print(f"pieces {pieces}: e2e step {1e3 * np.median(ts[2:]):.3f} ms (min {1e3 * min(ts[2:]):.3f})")
