"""TEST INFRASTRUCTURE. Golden text of the reference's --contin-debug / --epi-debug test functions
(ContingencyDebug, EpistasisDebug: algorithms/epistasis_func.cpp:84-103, 263-305), produced by running the UNMODIFIED
reference (oracle/_ref) in the build container on the first SNPs of the committed golden cohorts. Run from the repo
root: python tests/golden/make_golden_debug.py -> tests/golden/debug_prints.npz

EpistasisDebug builds its own CaseControlSet (cases = samples 0, 2, .., 398; controls = 1, 3, .., 399) and prints the
mask-on-the-fly tables to `out`; its log-likelihood lines go to stdout through printf and come from a stale C++
pairwise_epi_test that reads the 4x4 table as 3x3, so only the tables are pinned here."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
out = {}
for name, n_snps in (("cohort_missing", 20), ("cohort_complete", 14)):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    codes = g["codes"][:n_snps]
    per_level = {}
    for level in (4, 5):   # level 4 only for the comparison printed below
        R = oracle.Ref(n_snps, codes.shape[1], level)
        R.add_codes(codes)
        R.set_case_control(g["pheno"])
        per_level[level] = (R.run("ContingencyDebug"), R.run("EpistasisDebug"))
    # T5 is the canonical layout; T4's xx cells differ where the layouts pad differently (defects D5/D6)
    out[name + "_n_snps"] = n_snps
    out[name + "_contin_debug"] = per_level[5][0]
    out[name + "_epi_debug"] = per_level[5][1]
    print(name, len(per_level[5][0]), len(per_level[5][1]), "T4==T5:", per_level[4][0] == per_level[5][0], per_level[4][1] == per_level[5][1])
np.savez_compressed(os.path.join(GOLD, "debug_prints.npz"), **out)
