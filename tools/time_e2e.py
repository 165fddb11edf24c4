"""Breaks the marginal e2e step (bench.py `e2e`) into its pieces on configs[1]."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libgwaspp_b200 as gw  # noqa: E402

M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
h_counts = torch.empty((M, 8), dtype=torch.int32, pin_memory=True)
h_stats = torch.empty((M, 8), dtype=torch.float64, pin_memory=True)
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")


def best(f, n=10):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


st.select_case_control(case_mask=ca, ctrl_mask=co)
print(f"select_case_control (host masks)          {best(lambda: st.select_case_control(case_mask=ca, ctrl_mask=co)):7.3f} ms")
print(f"marginal_scan, device outputs             {best(lambda: st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)):7.3f} ms")
print(f"marginal_scan, pinned host outputs        {best(lambda: st.marginal_scan_into(0, M, counts=h_counts, stats=h_stats, on_device=False)):7.3f} ms")
print(f"D2H 16 MB counts (torch copy_)            {best(lambda: h_counts.copy_(d_counts, non_blocking=True)):7.3f} ms")
print(f"D2H 32 MB stats  (torch copy_)            {best(lambda: h_stats.copy_(d_stats, non_blocking=True)):7.3f} ms")
big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
big_h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
t = best(lambda: big_h.copy_(big_d, non_blocking=True), 5)
print(f"D2H 256 MB                                {t:7.3f} ms = {256 * 1.048576 / t:.1f} GB/s")
t = best(lambda: big_d.copy_(big_h, non_blocking=True), 5)
print(f"H2D 256 MB                                {t:7.3f} ms = {256 * 1.048576 / t:.1f} GB/s")
