"""K0 timing at configs[1]: the library prints the compaction kernel's CUDA-event time (GWASDEV_OPT_TRACE).
--tables forces the table-driven kernel used for cohorts beyond 32 768 samples (GWASDEV_OPT_SELECT_KERNEL)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401
import libgwaspp_b200 as gw  # noqa: E402

M, N, NCASE = 500_000, 10_000, 5_000
st = gw.GenoStore(M, N)
st.simulate(20121127)
st.set_select_mode(True)                     # eager: K0 inside select_case_control
st.set_option(gw.OPT_TRACE, 1)
if "--tables" in sys.argv:
    st.set_option(gw.OPT_SELECT_KERNEL, 1)
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    torch.cuda.synchronize()
    print("select_case_control (eager) ms", (time.perf_counter() - t0) * 1e3)
