"""Times the pair screen alone on configs[2] (or --snps/--samples): prints the screen kernel's CUDA-event time.
GWASDEV_MMA_DEBUG / GWASDEV_PAIR_ENGINE select kernel variants (diagnostics only; results are not checked here)."""
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
ap.add_argument("--lib", default="", help="alternative libgwasdev.so (A/B timing of kernel variants)")
a = ap.parse_args()
if a.lib:
    gw.LIB_PATH = os.path.abspath(a.lib)
ncase = a.cases or a.samples // 2
with gw.GenoStore(a.snps, a.samples) as st:
    st.simulate(20121127)
    st.select_case_control(gw.simulate_phenotype(20121127, a.samples, ncase))
    for r in range(a.reps):
        hits, s = st.pairwise_scan(30.0, shard=0, n_shards=a.shards)
        print(f"rep {r}: engine {s.engine} screen {s.screen_ms:.3f} ms total {s.total_ms:.3f} ms pairs {s.pairs_tested} "
              f"-> {s.pairs_tested / s.screen_ms / 1e6:.2f} G pairs/s, candidates {s.candidates}, hits {s.hits}", flush=True)
