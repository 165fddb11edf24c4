"""Shared test helpers (test infrastructure)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def planted_cohort(orc, seed, M, N, n_case, missing, n_planted):
    """Seeded synthetic cohort with a few case-only SNP-SNP dependencies so that the KSA screen crosses 30
    (same construction as tests/golden/make_golden.py)."""
    codes, pheno = orc.simulate(seed, M, N, n_case, missing_rate=missing)
    rng = np.random.default_rng(seed)
    cand = [r for r in range(M) if (codes[r] == 2).sum() > 0.04 * N]
    rng.shuffle(cand)
    for k in range(n_planted):
        i, j = sorted((cand[2 * k], cand[2 * k + 1]))
        src = codes[i].copy()
        src[src == 3] = 0
        codes[j, pheno == 1] = src[pheno == 1]
    return codes, pheno


def rel_close(a, b, rel):
    """element-wise |a-b| <= rel*|b|, NaN == NaN, inf == inf."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    same_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= rel * np.abs(b)
    return bool(np.all(ok | same_nan | same_inf))
