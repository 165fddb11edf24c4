"""Times the phases of the marginal e2e path (select_cc_maf call surface, host buffers) on configs[1]."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import libgwaspp_b200 as gw

M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
cm, tm = gw.stream_masks(pheno)
h_counts = torch.empty((M, 8), dtype=torch.int32, pin_memory=True)
h_stats = torch.empty((M, 8), dtype=torch.float64, pin_memory=True)
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
for r in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st.select_case_control(case_mask=cm, ctrl_mask=tm)
    t1 = time.perf_counter(); st.marginal_scan_into(0, M, counts=h_counts, stats=h_stats, on_device=False)
    t2 = time.perf_counter(); st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats); torch.cuda.synchronize()
    t3 = time.perf_counter(); h_counts.copy_(d_counts); h_stats.copy_(d_stats); torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"rep {r}: select {1e3*(t1-t0):.3f} ms | scan->host {1e3*(t2-t1):.3f} ms | scan->device {1e3*(t3-t2):.3f} ms (kernel {st.last_scan_ms():.3f}) | "
          f"torch D2H 48 MB {1e3*(t4-t3):.3f} ms", flush=True)
