import sys
sys.path.insert(0, "/root/repo")
import numpy as np, libgwaspp_b200 as gw
M, N, NC = 50000, 4000, 2000
with gw.GenoStore(M, N) as st:
    st.simulate(20121127, missing_rate=0.0002)
    st.select_case_control(gw.simulate_phenotype(20121127, N, NC))
    whole, s = st.pairwise_scan(30.0)
    st.set_pair_engine(1)
    popc, s1 = st.pairwise_scan(30.0)
    st.set_pair_engine(0)
    parts = [st.pairwise_scan(30.0, shard=k, n_shards=8) for k in range(8)]
    allp = np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"])
    print("whole", len(whole), "popc", len(popc), "equal", np.array_equal(whole, popc), "shards", [len(p[0]) for p in parts],
          "union equal", np.array_equal(allp, whole), "pairs", sum(p[1].pairs_tested for p in parts) == M * (M - 1) // 2,
          "ms per shard", [round(p[1].screen_ms, 2) for p in parts])
