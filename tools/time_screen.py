"""Times the pair screen alone on configs[2] (or --snps/--samples): prints the screen kernel's CUDA-event time.
--engine 1 times the AND+POPC engine, --lib another build of the library (A/B)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import libgwaspp_b200 as gw

ap = argparse.ArgumentParser()
ap.add_argument("--snps", type=int, default=50000)
ap.add_argument("--samples", type=int, default=4000)
ap.add_argument("--cases", type=int, default=0)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--shards", type=int, default=1)
ap.add_argument("--missing", type=float, default=0.0, help="per-genotype missing rate of the synthetic cohort")
ap.add_argument("--engine", type=int, default=0)
ap.add_argument("--classic-planes", action="store_true", help="two homozygote planes per SNP instead of the two rarest genotype classes")
ap.add_argument("--lib", default="", help="alternative libgwasdev.so (A/B timing of kernel variants)")
a = ap.parse_args()
if a.lib:
    gw.LIB_PATH = os.path.abspath(a.lib)
ncase = a.cases or a.samples // 2
with gw.GenoStore(a.snps, a.samples) as st:
    st.simulate(20121127, missing_rate=a.missing)
    st.select_case_control(gw.simulate_phenotype(20121127, a.samples, ncase))
    st.set_pair_engine(a.engine)
    if a.classic_planes:
        st.set_option(gw.OPT_CLASSIC_PLANES, 1)
    for r in range(a.reps):
        hits, s = st.pairwise_scan(30.0, shard=0, n_shards=a.shards)
        print(f"rep {r}: engine {s.engine} tiles {s.tiles} (9-cell {s.tiles_nine_cell}) screen {s.screen_ms:.3f} ms total {s.total_ms:.3f} ms pairs {s.pairs_tested} "
              f"-> {s.pairs_tested / s.screen_ms / 1e6:.2f} G pairs/s, candidates {s.candidates}, hits {s.hits}", flush=True)
    # phases of the computeBoost call surface with host buffers (what bench.py's pairwise e2e times)
    import time
    pheno = gw.simulate_phenotype(20121127, a.samples, ncase)
    cm, tm = gw.stream_masks(pheno)
    for r in range(3):
        t0 = time.perf_counter(); st.select_case_control(case_mask=cm, ctrl_mask=tm); st.synchronize()
        t1 = time.perf_counter(); hits, s = st.pairwise_scan(30.0, shard=0, n_shards=a.shards)
        t2 = time.perf_counter(); g = st.gtest(hits["i"], hits["j"]) if len(hits) else None
        t3 = time.perf_counter()
        print(f"e2e rep {r}: select {1e3 * (t1 - t0):.2f} ms, pairwise_scan {1e3 * (t2 - t1):.2f} ms (screen {s.screen_ms:.2f}), "
              f"gtest {1e3 * (t3 - t2):.2f} ms", flush=True)
