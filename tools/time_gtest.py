"""G-test alone on the hits of one screen (configs[3] by default; --cfg2): wall time of gwasdev_gtest with host buffers, several repeats.
GWASDEV_LIB=path selects an experiment build (build.py --variant NAME -DGWASDEV_GTEST_BLOCK=n)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import libgwaspp_b200 as gw  # noqa: E402

M, N, NC = (50000, 4000, 2000) if "--cfg2" in sys.argv else (500000, 10000, 5000)
st = gw.GenoStore(M, N)
st.simulate(20121127)
st.select_case_control(gw.simulate_phenotype(20121127, N, NC))
hits, s = st.pairwise_scan(30.0)
ref = None
for k in range(4):
    t0 = time.perf_counter()
    stat, z = st.gtest(hits["i"], hits["j"])
    t1 = time.perf_counter()
    if ref is None:
        ref = (stat.copy(), z.copy())
    same = np.array_equal(stat, ref[0]) and np.array_equal(z, ref[1], equal_nan=True)
    print(f"gtest rep {k}: {len(hits)} pairs {1e3 * (t1 - t0):.2f} ms, sum(stat) {float(np.sum(stat)):.6f}, repeatable {same}")
