"""Phase trace of the pairwise e2e step of bench.py (select -> pairwise_scan -> gtest) at configs[2] (or --cfg3: configs[3]); the library prints its phase times (GWASDEV_OPT_TRACE)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libgwaspp_b200 as gw  # noqa: E402

M, N, NC = (500000, 10000, 5000) if "--cfg3" in sys.argv else (50000, 4000, 2000)
st = gw.GenoStore(M, N)
st.set_option(gw.OPT_TRACE, 1)
st.simulate(20121127)
ph = gw.simulate_phenotype(20121127, N, NC)
cm, km = gw.stream_masks(ph)
for k in range(4):
    print(f"--- step {k}", file=sys.stderr)
    t0 = time.perf_counter()
    st.select_case_control(case_mask=cm, ctrl_mask=km)
    t1 = time.perf_counter()
    hits, s = st.pairwise_scan(30.0)
    t2 = time.perf_counter()
    st.gtest(hits["i"], hits["j"])
    t3 = time.perf_counter()
    print(f"step {k}: select {1e3 * (t1 - t0):.2f} ms, pairwise_scan {1e3 * (t2 - t1):.2f} ms, gtest {1e3 * (t3 - t2):.2f} ms", file=sys.stderr)
