import sys, time
sys.path.insert(0, '.')
import numpy as np, libgwaspp_b200 as gw
M,N,NC=50000,4000,2000
st=gw.GenoStore(M,N); st.simulate(20121127); ph=gw.simulate_phenotype(20121127,N,NC); st.select_case_control(ph)
for k in range(3):
    t=time.perf_counter(); hits,s=st.pairwise_scan(30.0); print("scan",k,time.perf_counter()-t, s.screen_ms, s.total_ms, len(hits), file=sys.stderr)
# marginal e2e pieces
M,N,NC=500000,10000,5000
st2=gw.GenoStore(M,N); st2.simulate(20121127); ph=gw.simulate_phenotype(20121127,N,NC)
cm,km=gw.stream_masks(ph)
for k in range(4):
    t=time.perf_counter(); st2.select_case_control(case_mask=cm,ctrl_mask=km); t1=time.perf_counter(); o=st2.marginal_scan(mi=False); t2=time.perf_counter()
    print("marginal e2e: select %.2f ms, scan+D2H(pageable) %.2f ms" % ((t1-t)*1e3,(t2-t1)*1e3), file=sys.stderr)
