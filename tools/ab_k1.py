"""A/B of the compacted marginal scan (K1) between two builds of the library, same process order, configs[1].
usage: python tools/ab_k1.py libA.so libB.so   (plain ctypes: works with round-1 builds too)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libgwaspp_b200.maf_spectrum import MAF_SPECTRUM  # noqa: E402

M, N, NCASE = 500_000, 10_000, 5_000


def run(path):
    L = C.CDLL(os.path.abspath(path))
    L.gwasdev_last_scan_ms.restype = C.c_double
    L.gwasdev_last_scan_ms.argtypes = [C.c_void_p]
    L.gwasdev_plane_blocks.restype = C.c_uint32
    h = C.c_void_p()
    assert L.gwasdev_create(C.c_uint64(M), C.c_uint32(N), 0, C.byref(h)) == 0
    bins = np.asarray(MAF_SPECTRUM["affy6"], np.uint32)
    assert L.gwasdev_simulate(h, C.c_uint64(20121127), bins.ctypes.data_as(C.c_void_p), C.c_uint32(0)) == 0
    pheno = np.zeros(N, np.uint8)
    L.gwasdev_simulate_phenotype(C.c_uint64(20121127), C.c_uint32(N), C.c_uint32(NCASE), pheno.ctypes.data_as(C.c_void_p))
    P = L.gwasdev_plane_blocks(C.c_uint32(N))
    ca, co = np.zeros(P * 16, np.uint8), np.zeros(P * 16, np.uint8)
    ca[:N], co[:N] = pheno == 1, pheno == 0
    w = (1 << np.arange(16)).astype(np.uint32)
    cm = (ca.reshape(P, 16) * w).sum(1).astype(np.uint16)
    km = (co.reshape(P, 16) * w).sum(1).astype(np.uint16)
    assert L.gwasdev_set_select_mode(h, 1) == 0
    assert L.gwasdev_select_case_control(h, cm.ctypes.data_as(C.c_void_p), km.ctypes.data_as(C.c_void_p)) == 0
    import torch
    d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
    d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
    ms = []
    for _ in range(30):
        assert L.gwasdev_marginal_scan(h, C.c_uint64(0), C.c_uint64(M), C.c_void_p(d_counts.data_ptr()), None, C.c_void_p(d_stats.data_ptr()), 1) == 0
        ms.append(L.gwasdev_last_scan_ms(h))
    ms_c = []
    for _ in range(30):       # counts only
        assert L.gwasdev_marginal_scan(h, C.c_uint64(0), C.c_uint64(M), C.c_void_p(d_counts.data_ptr()), None, None, 1) == 0
        ms_c.append(L.gwasdev_last_scan_ms(h))
    L.gwasdev_destroy(h)
    return float(np.median(ms[5:])), float(np.min(ms[5:])), float(np.median(ms_c[5:]))


for rnd in range(2):
    for path in sys.argv[1:]:
        med, mn, cnt = run(path)
        print(f"{os.path.basename(path):28s} K1 counts+stats median {med:.4f} ms (min {mn:.4f}) = {M * N / 4 / med / 1e6:7.1f} GB/s; counts only {cnt:.4f} ms", flush=True)
