"""The real binding, compiled and run: class DeviceGenotypeTable : public GenoTable (libgwaspp_b200/binding/), built against the
reference's own headers, plugged into GeneticData's factory as compression level 6 (oracle/ref_build/Makefile patches the two
factory lines at build time) and driven by the reference's UNMODIFIED compute() and test functions -- inline_maf_print,
select_cc_maf, inline_cc_maf, computeMargins, computeBoost, computeGTest, ContingencyDebug, EpistasisDebug. Every result is
compared with the same functions running on the reference's own level-5 table (oracle/_ref/libgwasref.so) in this process.

The libraries are built in the build container (they need /root/reference) and travel to the GPU box; nothing here reads
/root/reference.
"""
import re

import numpy as np
import pytest

import oracle
from helpers import planted_cohort, rel_close

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (oracle.have_ref() and oracle.have_ref_dev()), reason="reference builds absent")]

LAPSE = re.compile(r"\d+\.\d{9}s")


def pair(orc, codes, pheno):
    M, N = codes.shape
    ref, dev = oracle.Ref(M, N, 5), oracle.RefDev(M, N, 6)
    for r in (ref, dev):
        r.add_codes(codes)
        r.set_case_control(pheno)
    return ref, dev


@pytest.mark.parametrize("miss", [0.0, 0.03])
def test_reference_functions_on_the_device_table(orc, miss):
    M, N, NCASE = 260, 700, 330
    codes, pheno = planted_cohort(orc, 4242, M, N, NCASE, miss, 6)
    pheno = pheno.copy()
    pheno[::41] = 2                                             # some samples in neither class
    ref, dev = pair(orc, codes, pheno)
    assert (dev.n_snps, dev.n_samples, dev.n_cases, dev.n_controls) == (ref.n_snps, ref.n_samples, ref.n_cases, ref.n_controls)
    # codec and cell access through the table's own virtuals
    for r in range(0, M, 37):
        for c in range(0, N, 53):
            assert dev.call_at(r, c) == ref.call_at(r, c)
    # the reference's own test functions, unmodified, through compute()
    assert dev.run("inline_maf_print") == ref.run("inline_maf_print")
    for fn in ("select_cc_maf", "genotype_dist_performance"):
        a, b = dev.run(fn), ref.run(fn)
        assert LAPSE.sub("T", a) == LAPSE.sub("T", b)           # same lines; the lapse values differ, of course
    # inline_cc_maf copies the CaseControlSet by value (algorithms/maf_func.cpp:300: no copy constructor, so the copy's
    # destructor frees the original's mask buffers): nothing may use that GeneticData's set afterwards -- separate objects
    ref2, dev2 = pair(orc, codes, pheno)
    assert LAPSE.sub("T", dev2.run("inline_cc_maf")) == LAPSE.sub("T", ref2.run("inline_cc_maf"))
    # per-row overloads: whole cohort (aa / ab / bb; the reference's xx counts its row padding), mask-on-the-fly, pre-selected
    ref.select()
    dev.select()
    for r in range(M):
        assert np.array_equal(dev.dist(r)[:3], ref.dist(r)[:3])
        assert dev.dist(r)[3] == N - ref.dist(r)[:3].sum()
        for mode in (0, 1):
            assert np.array_equal(dev.cc_dist(r, mode)[0], ref.cc_dist(r, mode)[0]), (r, mode)
        (cd, md), (cr, mr) = dev.cc_dist(r, 2), ref.cc_dist(r, 2)
        assert np.array_equal(cd, cr)
        for f in ("margins", "cases", "controls", "pbc", "pca"):
            assert np.array_equal(md[f], mr[f])
        assert rel_close(md["entropy"], mr["entropy"], 1e-12) and rel_close(md["entropy_y"], mr["entropy_y"], 1e-12)
    # computeMargins
    md, mr = dev.margins(), ref.margins()
    for f in ("margins", "cases", "controls", "pbc", "pca"):
        assert np.array_equal(md[f], mr[f])
    assert rel_close(md["entropy"], mr["entropy"], 1e-12) and rel_close(md["entropy_y"], mr["entropy_y"], 1e-12)
    # per-pair overloads, walked like the reference's loops and at random
    rng = np.random.default_rng(1)
    pairs = [(i, j) for i in range(0, 12) for j in range(i + 1, 90)] + [tuple(sorted(rng.choice(M, 2, replace=False))) for _ in range(200)]
    for mode in (0, 1, 2, 3):
        for i, j in pairs:
            (a0, a1), (b0, b1) = dev.pair_table(int(i), int(j), mode), ref.pair_table(int(i), int(j), mode)
            assert np.array_equal(a0, b0) and np.array_equal(a1, b1), (mode, i, j)
    # computeBoost: pre-screen over every pair + computeGTest + the printed result lines
    td, tr = dev.run("computeBoost"), ref.run("computeBoost")
    (hd, ld), (hr, lr) = oracle.parse_boost_output(td), oracle.parse_boost_output(tr)
    assert ld == lr and len(hr) >= 3 and [h[:2] for h in hd] == [h[:2] for h in hr]
    assert rel_close([h[2] for h in hd], [h[2] for h in hr], 1e-5) and rel_close([h[3] for h in hd], [h[3] for h in hr], 1e-5)   # "%f" prints 6 decimals
    assert [ln for ln in td.splitlines() if not LAPSE.fullmatch(ln.strip())] == [ln for ln in tr.splitlines() if not LAPSE.fullmatch(ln.strip())]
    pi, pj = np.array([h[0] for h in hr], np.uint32), np.array([h[1] for h in hr], np.uint32)
    (sd, zd), (sr, zr) = dev.gtest(pi, pj), ref.gtest(pi, pj)
    assert rel_close(sd, sr, 1e-11) and rel_close(zd, zr, 1e-12)


def test_debug_printers_on_the_device_table(orc):
    """ContingencyDebug / EpistasisDebug (epistasis_func.cpp:84-103, 263-305): every pair's 4x4 tables as text."""
    M, N = 36, 420
    codes, pheno = orc.simulate(99, M, N, 200, missing_rate=0.02)
    ref, dev = pair(orc, codes, pheno)
    assert dev.run("ContingencyDebug") == ref.run("ContingencyDebug")
    assert dev.run("EpistasisDebug") == ref.run("EpistasisDebug")


def test_reselection_and_row_updates_through_the_reference_api(orc):
    """selectCaseControl with another set, then rows rewritten through addGenotypeRow: the blocks the binding serves must follow."""
    M, N = 120, 300
    codes, pheno = orc.simulate(5, M, N, 140, missing_rate=0.01)
    ref, dev = pair(orc, codes, pheno)
    for r in (ref, dev):
        r.select()
    assert all(np.array_equal(dev.cc_dist(r, 1)[0], ref.cc_dist(r, 1)[0]) for r in range(M))
    ph2 = 1 - pheno
    for r in (ref, dev):
        r.set_case_control(ph2)
        r.select()
    assert all(np.array_equal(dev.cc_dist(r, 1)[0], ref.cc_dist(r, 1)[0]) for r in range(M))
    codes2, _ = orc.simulate(6, 10, N, 140)
    for r in (ref, dev):
        r.add_codes(codes2, first_row=50)
        r.select()
    assert all(np.array_equal(dev.cc_dist(r, 1)[0], ref.cc_dist(r, 1)[0]) for r in range(M))
    assert np.array_equal(dev.pair_table(49, 55, 2)[0], ref.pair_table(49, 55, 2)[0])
