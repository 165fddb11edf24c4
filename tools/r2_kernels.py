"""Round-2 kernel timings at configs[1] (500 000 SNPs x 5 000 / 5 000 samples), CUDA events inside the library:
K0 (compaction, eager selection), K1 (compacted scan), K1' first scan of a table (MODE 1) and re-selection scans (MODE 2),
and the host-output forms of the select_cc_maf surface."""
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
d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
GB = M * N / 4 / 1e9


def med(f, n=8):
    out = []
    for _ in range(n):
        out.append(f())
    return float(np.median(out[2:]))


def k0():
    st.set_select_mode(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    st.synchronize()
    return (time.perf_counter() - t0) * 1e3


st.set_option(gw.OPT_TRACE, 1)         # the library prints the compaction kernel's CUDA-event time on stderr
t = med(k0)
st.set_option(gw.OPT_TRACE, 0)
print(f"eager select (host tables + K0), wall: {t:.3f} ms; K0 traffic 2 x {GB:.2f} GB")


def scan():
    st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
    return st.last_scan_ms()


t = med(scan)
print(f"K1 compacted scan:            {t:.4f} ms = {GB / t * 1e3:7.1f} GB/s")
st.set_select_mode(False)
st.set_option(gw.OPT_ROW_TOTALS, 1)


def first():
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    return scan()


t = med(first)
print(f"K1' MODE 1 (no cached totals): {t:.4f} ms = {GB / t * 1e3:7.1f} GB/s")
st.set_option(gw.OPT_ROW_TOTALS, 0)
first()
t = med(first)
print(f"K1' MODE 2 (cached totals):    {t:.4f} ms = {GB / t * 1e3:7.1f} GB/s")
h_compact = torch.empty((M, 4), dtype=torch.int64, pin_memory=True)
h_counts = torch.empty((M, 8), dtype=torch.int32, pin_memory=True)
h_stats = torch.empty((M, 8), dtype=torch.float64, pin_memory=True)


def wall(f):
    def g():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.select_case_control(case_mask=ca, ctrl_mask=co)
        f()
        return (time.perf_counter() - t0) * 1e3
    return g


for pieces in (0, 1, 2, 4, 8):
    st.set_option(gw.OPT_SCAN_PIECES, pieces)
    a = med(wall(lambda: st.marginal_scan_compact_into(0, M, records=h_compact, on_device=False)))
    b = med(wall(lambda: st.marginal_scan_into(0, M, counts=h_counts, stats=h_stats, on_device=False)))
    print(f"e2e select + scan to host, pieces {pieces}: compact (32 B/SNP) {a:.3f} ms, full (96 B/SNP) {b:.3f} ms")
