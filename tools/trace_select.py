"""K0 timing at configs[1]: run with GWASDEV_TRACE=1 (prints the compaction kernel's CUDA-event time).
GWASDEV_SELECT_TABLE_KERNEL=1 forces the table-driven kernel used for cohorts beyond 32 768 samples."""
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
pheno = gw.simulate_phenotype(20121127, N, NCASE)
ca, co = gw.stream_masks(pheno)
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st.select_case_control(case_mask=ca, ctrl_mask=co)
    torch.cuda.synchronize()
    print("select_case_control (eager) ms", (time.perf_counter() - t0) * 1e3)
