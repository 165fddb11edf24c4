"""Regenerates tests/golden/*.npz and the perl fixtures FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, perl, and oracle/_ref built by
oracle/ref_build/Makefile):   python tests/golden/make_golden.py

Sources of truth captured here:
  (i)   scripts/perl/genotype_set_builder.pl  -> *.tped/*.tfam + *.expected*.dist (xx aa ab bb per marker)
  (ii)  the reference's own tables/test functions run through oracle/_ref (T5, cross-checked against T4)
  (iii) src/test/pairwise.c:19-37 five literature 3x3x2 tables -> ll and pchisq(ll, 4)
The perl script uses unseeded rand(); its output is committed together with its expectation files.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PERL = "/root/reference/scripts/perl/genotype_set_builder.pl"

KATS = {   # src/test/pairwise.c:19-37
    "boost0": ([5, 24, 14, 57, 138, 96, 151, 315, 187], [4, 16, 10, 71, 176, 86, 120, 330, 200]),
    "boost1": ([227, 225, 29, 191, 189, 58, 24, 33, 11], [229, 219, 67, 195, 199, 26, 33, 42, 3]),
    "boost2": ([32, 103, 93, 49, 250, 203, 52, 94, 111], [33, 121, 96, 76, 203, 227, 20, 124, 113]),
    "biforce1": ([2, 10, 12, 31, 130, 87, 75, 295, 345], [3, 14, 3, 16, 90, 143, 84, 353, 307]),
    "biforce2": ([21, 214, 289, 7, 119, 260, 0, 0, 77], [12, 123, 542, 3, 69, 223, 4, 14, 23]),
}


def planted_cohort(seed, M, N, n_case, missing, n_planted):
    """Seeded cohort (oracle generator) with a few case-only SNP-SNP dependencies planted so that
    computeBoost's threshold of 30 (epistasis_func.cpp:390) is crossed by some pairs."""
    O = oracle.Oracle()
    codes, pheno = O.simulate(seed, M, N, n_case, missing_rate=missing)
    rng = np.random.default_rng(seed)
    cand = [r for r in range(M) if (codes[r] == 2).sum() > 0.04 * N]
    rng.shuffle(cand)
    for k in range(n_planted):
        i, j = sorted((cand[2 * k], cand[2 * k + 1]))
        src = codes[i].copy()
        src[src == 3] = 0
        codes[j, pheno == 1] = src[pheno == 1]
    return codes, pheno


def capture_cohort(name, seed, M, N, n_case, missing, n_planted):
    codes, pheno = planted_cohort(seed, M, N, n_case, missing, n_planted)
    out = {"codes": codes, "pheno": pheno}
    per_level = {}
    for level in (5, 4):
        R = oracle.Ref(M, N, level)
        R.add_codes(codes)
        R.set_case_control(pheno)
        d = {}
        d["whole"] = np.stack([R.dist(r) for r in range(M)])
        d["cc_masked"] = np.stack([R.cc_dist(r, 0)[0] for r in range(M)])
        d["inline_maf_print"] = R.run("inline_maf_print")
        R.select()
        d["cc_selected"] = np.stack([R.cc_dist(r, 1)[0] for r in range(M)])
        d["margins"] = R.margins()
        rng = np.random.default_rng(seed + 1)
        pairs = [(int(a), int(b)) for a, b in (sorted(rng.choice(M, 2, replace=False)) for _ in range(48))]
        d["pairs"] = np.array(pairs, np.uint32)
        for mode in (0, 1, 2, 3):
            t = [R.pair_table(i, j, mode) for i, j in pairs]
            d[f"tables_mode{mode}"] = np.stack([np.concatenate(x) for x in t])
        d["boost_text"] = R.run("computeBoost")
        hits, located = oracle.parse_boost_output(d["boost_text"])
        d["boost_located"] = located
        d["boost_hits"] = np.array([(h[0], h[1]) for h in hits], np.uint32).reshape(-1, 2)
        d["boost_exact"] = np.array([h[2] for h in hits])
        d["boost_z"] = np.array([h[3] for h in hits])
        gp = np.array(pairs[:16] + [tuple(h) for h in d["boost_hits"].tolist()], np.uint32)
        d["gtest_pairs"] = gp
        d["gtest_stat"], d["gtest_z"] = R.gtest(gp[:, 0], gp[:, 1])
        if level == 5:
            d["raw_rows"] = np.stack([R.raw_row(r) for r in range(M)])
            d["sel_rows"] = np.stack([R.selected_row(r)[0] for r in range(M)])
            d["n_case"], d["n_ctrl"] = R.n_cases, R.n_controls
        per_level[level] = d
    # cross-layout identity T4 == T5 on everything both implement correctly
    a, b = per_level[5], per_level[4]
    assert np.array_equal(a["whole"][:, :3], b["whole"][:, :3])
    assert a["margins"].tobytes() == b["margins"].tobytes()
    assert np.array_equal(a["tables_mode3"], b["tables_mode3"])
    assert np.array_equal(a["boost_hits"], b["boost_hits"]) and np.array_equal(a["boost_exact"], b["boost_exact"])
    out.update(a)
    # the KSA screen value itself is not printed by the reference; what it prints is the number of
    # pairs whose screen value exceeded 30 ("Located N potential interactions") and, per printed hit,
    # the exact G-test statistic and z.
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    print(name, "hits:", len(a["boost_hits"]), "located:", a["boost_located"])


def perl_fixture(name, markers, individs, cc):
    tmp = os.path.join(GOLD, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    args = ["perl", PERL, "--markers", str(markers), "--individs", str(individs), "--tplink", name]
    if cc:
        args.append("--case-control")
    subprocess.check_call(args, cwd=tmp, stdout=subprocess.DEVNULL)
    for fn in os.listdir(tmp):
        src = os.path.join(tmp, fn)
        with open(src) as f:
            txt = f.read()
        if fn.endswith((".tped", ".tfam")):
            txt = txt.replace(" ", "\t")   # the harness parses with '\t' (src/test/gwas_basic.cpp:145)
        with open(os.path.join(GOLD, "perl_" + fn), "w") as f:
            f.write(txt)
        os.remove(src)
    os.rmdir(tmp)
    # what the reference itself prints for these files, all three hot layouts
    outs = {}
    for level in (3, 4, 5):
        R = oracle.Ref(tped=os.path.join(GOLD, f"perl_{name}.tped"), tfam=os.path.join(GOLD, f"perl_{name}.tfam"),
                       level=level)
        outs[level] = R.run("inline_maf_print")
        if cc and level == 5:
            R.select()
            cc_counts = np.stack([R.cc_dist(r, 1)[0] for r in range(R.n_snps)])
            np.save(os.path.join(GOLD, f"perl_{name}.ref_cc_counts.npy"), cc_counts)
    assert outs[3] == outs[4] == outs[5]
    with open(os.path.join(GOLD, f"perl_{name}.ref_inline_maf_print.txt"), "w") as f:
        f.write(outs[5])


def kats():
    L = oracle.Ref.lib()
    import ctypes as C
    rows = []
    for k, (cs, ct) in KATS.items():
        p = C.c_double()
        a = (C.c_int * 9)(*cs)
        b = (C.c_int * 9)(*ct)
        ll = L.gwasref_pairwise_c(a, b, C.byref(p))
        rows.append((k, cs, ct, ll, p.value))
    np.savez(os.path.join(GOLD, "pairwise_c_kats.npz"),
             names=np.array([r[0] for r in rows]), cs=np.array([r[1] for r in rows], np.int32),
             ct=np.array([r[2] for r in rows], np.int32), ll=np.array([r[3] for r in rows]),
             p=np.array([r[4] for r in rows]))
    for r in rows:
        print(r[0], repr(r[3]), repr(r[4]))


if __name__ == "__main__":
    oracle.build_ref()
    kats()
    perl_fixture("simple", 12, 150, False)
    perl_fixture("cc", 10, 130, True)
    capture_cohort("cohort_missing", 20121127, 96, 500, 257, 0.02, 6)
    capture_cohort("cohort_complete", 20121128, 64, 400, 200, 0.0, 5)
