"""GPU parity tests: the CUDA path, called through the C-ABI, against the pinned oracle and the golden
vectors captured from the unmodified reference.

Bars (BASELINE.json north_star): genotype counts, contingency tables, layouts and hit sets bit-exact;
fp64 statistics within 1e-12 relative (written at each assert); p-values within 1e-10 relative.
Nothing here reads /root/reference.
"""
import numpy as np
import pytest

import libgwaspp_b200 as gw
from helpers import load_golden, planted_cohort, rel_close

pytestmark = pytest.mark.gpu

REL_F64 = 1e-12     # fp64 statistic tolerance (north_star)
COHORTS = ["cohort_missing", "cohort_complete"]


def make_store(orc, codes, pheno=None):
    M, N = codes.shape
    st = gw.GenoStore(M, N)
    st.put_rows(orc.pack_codes(codes))
    if pheno is not None:
        st.select_case_control(pheno)
    return st


@pytest.mark.parametrize("k0", ["columns", "tables"])
@pytest.mark.parametrize("name", COHORTS)
def test_store_layout_roundtrip(orc, name, k0):
    g = load_golden(name)
    codes, pheno = g["codes"], g["pheno"]
    with make_store(orc, codes) as st:
        if k0 == "tables":      # the table-driven compaction kernel (cohorts beyond 32 768 samples) on the same fixtures
            st.set_option(gw.OPT_SELECT_KERNEL, 1)
        assert np.array_equal(st.get_rows(), g["raw_rows"])                    # a5 layout incl. headers
        assert np.array_equal(st.get_rows(7, 5), g["raw_rows"][7:12])
        txt = {0: "AA", 1: "AC", 2: "CC", 3: "00"}
        for r in range(0, codes.shape[0], 9):
            for c in range(0, codes.shape[1], 37):
                assert st.call_at(r, c) == txt[int(codes[r, c])]               # operator() + decodeGenotype
        st.select_case_control(pheno)
        assert (st.n_case, st.n_ctrl) == (int(g["n_case"]), int(g["n_ctrl"]))
        assert np.array_equal(st.get_selected_rows(), g["sel_rows"])           # a8: K0 on the device
        # re-selection with swapped classes keeps working on the resident raw store
        st.select_case_control(1 - pheno.astype(np.int64))
        sel2, nca2, nco2 = orc.select(g["raw_rows"], codes.shape[1], (1 - pheno).astype(np.uint8))
        assert np.array_equal(st.get_selected_rows(), sel2)


@pytest.mark.parametrize("name", COHORTS)
def test_marginal_scan_matches_reference_golden(orc, name):
    g = load_golden(name)
    codes, pheno = g["codes"], g["pheno"]
    with make_store(orc, codes, pheno) as st:
        out = st.marginal_scan()
        assert np.array_equal(out["counts"], g["cc_selected"])                 # a9, bit-exact
        mi, ref = out["mi"], g["margins"]
        for f in ("margins", "cases", "controls"):
            assert np.array_equal(mi[f], ref[f])
        assert np.array_equal(mi["dPbc"], ref["pbc"]) and np.array_equal(mi["dPca"], ref["pca"])   # pure divisions
        assert rel_close(mi["dMarginalEntropy"], ref["entropy"], REL_F64)
        assert rel_close(mi["dMarginalEntropy_Y"], ref["entropy_y"], REL_F64)
        for r in range(codes.shape[0]):
            ca, co = out["counts"][r, :4], out["counts"][r, 4:]
            s = out["stats"][r]
            assert rel_close(s["maf_ref_case"], orc.maf_reference(ca)[0], 1e-15)    # a11
            assert rel_close(s["maf_ref_ctrl"], orc.maf_reference(co)[0], 1e-15)
            x, p = orc.chi2_allelic(ca, co)
            assert rel_close(s["chi2_allelic"], x, REL_F64) and rel_close(s["p_allelic"], p, 1e-10)
            x, p, df = orc.chi2_genotypic(ca, co)
            assert s["df_genotypic"] == df
            assert rel_close(s["chi2_genotypic"], x, REL_F64) and rel_close(s["p_genotypic"], p, 1e-10)
        # sub-range call and the other SingleMarkerAnalyzable overloads
        part = st.marginal_scan(10, 33, mi=False, stats=False)
        assert np.array_equal(part["counts"], g["cc_selected"][10:33])
        assert np.array_equal(st.counts(2), g["cc_selected"])
        assert np.array_equal(st.counts(1), g["cc_masked"])
        whole = st.counts(0)
        assert np.array_equal(whole[:, :3], g["whole"][:, :3])
        assert np.array_equal(whole[:, 3], codes.shape[1] - g["whole"][:, :3].sum(1))   # what inline_maf_print prints


@pytest.mark.parametrize("N,lanes", [(677, 0), (2000, 16), (3900, 32), (130, 8), (1111, 16)])
def test_fused_select_scan_equals_compacted_scan(orc, N, lanes):
    """The first scan after a selection counts through the masks on the raw rows (no compaction, the reference's
    mask-on-the-fly overload compressed_genotype_table5.cpp:609-657); later scans stream the compacted rows (:703-747).
    Both must give the oracle's counts and bit-identical statistics -- also when samples belong to neither class or are
    flagged in both masks (a case for the compaction, :541-561), and for every lane-group width of the kernels."""
    M = 301
    codes, _ = orc.simulate(77 + N, M, N, N // 2, missing_rate=0.03)
    rng = np.random.default_rng(N)
    pheno = rng.choice([0, 1, 2], size=N, p=[0.45, 0.4, 0.15]).astype(np.uint8)          # 2 = neither
    rows = orc.pack_codes(codes)
    ca, co = gw.stream_masks(pheno)
    both = rng.choice(N, 9, replace=False)
    for c in both:                                                                        # flagged in both masks
        ca[c >> 4] |= np.uint16(1 << (c & 15))
        co[c >> 4] |= np.uint16(1 << (c & 15))
    eff = pheno.copy()
    eff[both] = 1
    sel, nca, nco = orc.select(rows, N, eff)
    want = orc.cc_counts_selected(sel, nca, nco)
    with gw.GenoStore(M, N) as lazy, gw.GenoStore(M, N) as eager:
        eager.set_select_mode(True)
        outs = []
        for st in (lazy, eager):
            st.set_option(gw.OPT_LANES_PER_ROW, lanes)            # lanes cooperating on one row (0: chosen from the row length)
            st.put_rows(rows)
            st.select_case_control(case_mask=ca, ctrl_mask=co)
            assert (st.n_case, st.n_ctrl) == (nca, nco)
            outs.append(st.marginal_scan())
        second = lazy.marginal_scan()                                                     # compacted rows, built on demand
        for o in outs + [second]:
            assert np.array_equal(o["counts"], want)
            for k in ("mi", "stats"):
                assert o[k].tobytes() == outs[1][k].tobytes()
        assert np.array_equal(lazy.get_selected_rows(), sel) and np.array_equal(eager.get_selected_rows(), sel)
        # sub-range as the first scan after a new selection (masked kernel with a row offset)
        lazy.select_case_control(case_mask=ca, ctrl_mask=co)
        sub = lazy.marginal_scan(17, 230)
        assert np.array_equal(sub["counts"], want[17:230]) and sub["stats"].tobytes() == outs[1]["stats"][17:230].tobytes()


@pytest.mark.parametrize("name", COHORTS)
def test_pair_tables_all_overloads(orc, name):
    g = load_golden(name)
    with make_store(orc, g["codes"], g["pheno"]) as st:
        pi, pj = g["pairs"][:, 0], g["pairs"][:, 1]
        for mode in (0, 1, 2, 3):
            assert np.array_equal(st.pair_tables(pi, pj, mode), g[f"tables_mode{mode}"]), mode


@pytest.mark.parametrize("name", COHORTS)
def test_ksa_gtest_and_screen_match_reference_golden(orc, name):
    g = load_golden(name)
    codes, pheno = g["codes"], g["pheno"]
    M, N = codes.shape
    with make_store(orc, codes, pheno) as st:
        sel, nca, nco = g["sel_rows"], int(g["n_case"]), int(g["n_ctrl"])
        mar = orc.margins(sel, nca, nco)
        # every pair, fp64 KSA vs the oracle (NaN pattern included)
        ii, jj = np.triu_indices(M, 1)
        dev = st.ksa(ii, jj)
        want = np.array([orc.ksa(*orc.pair_table(3, int(i), int(j), sel=sel, nca=nca, nco=nco, mar=mar), mar[i], mar[j],
                                 nca + nco) for i, j in zip(ii, jj)])
        assert np.array_equal(np.isnan(dev), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.all(np.abs(dev[ok] - want[ok]) <= REL_F64 * np.maximum(np.abs(want[ok]), 1.0))
        # exhaustive screen == computeBoost's "Located N potential interactions"
        hits, stats = st.pairwise_scan(30.0)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, nca, nco, 30.0)
        assert len(hits) == int(g["boost_located"]) == len(hi)
        assert np.array_equal(hits["i"], hi) and np.array_equal(hits["j"], hj)
        assert rel_close(hits["stat"], hs, REL_F64)
        assert stats.pairs_tested == M * (M - 1) // 2 and stats.hits == len(hits)
        # exact G-test + z (computeGTest) on the reference's own pair list
        gp = g["gtest_pairs"]
        s, z = st.gtest(gp[:, 0], gp[:, 1])
        assert rel_close(s, g["gtest_stat"], 1e-11)     # IPF stops on an absolute 1e-3 criterion; see DESIGN.md
        assert rel_close(z, g["gtest_z"], REL_F64)
        # what computeBoost prints: pairs whose exact statistic exceeds 30 (epistasis_func.cpp:497-505)
        es, ez = st.gtest(hits["i"], hits["j"])
        kept = es > 30.0
        assert np.array_equal(np.stack([hits["i"][kept], hits["j"][kept]], 1), g["boost_hits"])
        assert np.allclose(es[kept], g["boost_exact"], rtol=0, atol=5.1e-7)    # printed with %f
        assert np.allclose(ez[kept], g["boost_z"], rtol=0, atol=5.1e-7)


def test_pairwise_c_known_answers(orc):
    k = load_golden("pairwise_c_kats")
    ll, p = gw.pairwise_epi_test(k["cs"], k["ct"])
    assert rel_close(ll, k["ll"], REL_F64) and rel_close(p, k["p"], 1e-10)


@pytest.mark.parametrize("seed,M,N,ncase,miss,planted", [
    (11, 700, 1200, 600, 0.0, 12),      # no missing data: 4-corner tiles only, several 64x64 tiles + ragged edge
    (12, 333, 900, 371, 0.01, 8),       # missing everywhere: 9-cell tiles
    (13, 200, 97, 31, 0.0, 4),          # tiny, class sizes below one word / ragged
    (14, 130, 2048, 1024, 0.0, 4),      # class size an exact multiple of 32 and 64
])
def test_exhaustive_screen_against_oracle(orc, seed, M, N, ncase, miss, planted):
    codes, pheno = planted_cohort(orc, seed, M, N, ncase, miss, planted)
    if miss > 0:                                        # leave some tiles completely called -> mixed kernels
        codes[:128][codes[:128] == 3] = 0
    with make_store(orc, codes, pheno) as st:
        sel = st.get_selected_rows()
        nca, nco = st.n_case, st.n_ctrl
        mar = orc.margins(sel, nca, nco)
        hi, hj, hs, stt = orc.boost_screen(sel, mar, nca, nco, 30.0)
        hits, stats = st.pairwise_scan(30.0)
        assert len(hi) >= planted // 2                  # the test is not vacuous
        assert np.array_equal(hits["i"], hi) and np.array_equal(hits["j"], hj)
        assert rel_close(hits["stat"], hs, REL_F64)
        assert stats.pairs_tested == M * (M - 1) // 2
        assert stats.candidates >= stats.hits
        # sharded over 3 "GPUs": disjoint tile sets whose union is the whole result
        parts = [st.pairwise_scan(30.0, shard=k, n_shards=3) for k in range(3)]
        assert sum(p[1].pairs_tested for p in parts) == M * (M - 1) // 2
        merged = np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"])
        assert np.array_equal(merged, hits)
        # a lower threshold exercises the candidate path with many more survivors
        hits5, _ = st.pairwise_scan(12.0)
        hi5, hj5, hs5, _ = orc.boost_screen(sel, mar, nca, nco, 12.0)
        assert np.array_equal(hits5["i"], hi5) and np.array_equal(hits5["j"], hj5) and rel_close(hits5["stat"], hs5, REL_F64)
        # fp32 screen epilogue stays well inside its safety margin
        ii, jj = np.triu_indices(min(M, 150), 1)
        f32, f64 = st.ksa_screen_f32(ii, jj), st.ksa(ii, jj)
        ok = ~np.isnan(f64)
        assert np.array_equal(np.isnan(f32), np.isnan(f64))
        assert np.max(np.abs(f32[ok] - f64[ok])) < 0.125


def test_device_generator_matches_host_restatement(orc):
    M, N = 257, 1031
    for miss in (0.0, 0.03):
        codes, _ = orc.simulate(20121127, M, N, 500, missing_rate=miss)
        with gw.GenoStore(M, N) as st:
            st.simulate(20121127, missing_rate=miss)
            assert np.array_equal(st.get_rows(), orc.pack_codes(codes))


def test_error_behaviour():
    with gw.GenoStore(10, 40) as st:
        with pytest.raises(gw.GwasDevError, match="select_case_control"):
            st.marginal_scan()
        with pytest.raises(gw.GwasDevError):
            st.put_rows(np.zeros((11, 2 * st.P + 1), np.uint16))
        with pytest.raises(gw.GwasDevError, match="both masks are empty"):
            st.select_case_control(np.full(40, 2))
        st.select_case_control(np.arange(40) % 2)
        with pytest.raises(gw.GwasDevError, match="outside the table"):
            st.pair_tables([0], [10], 3)
        # an all-missing table: every statistic is NaN, nothing is a hit, nothing crashes
        out = st.marginal_scan()
        assert np.all(out["counts"][:, 3] == 20) and np.all(out["counts"][:, :3] == 0)
        hits, stats = st.pairwise_scan(30.0)
        assert len(hits) == 0 and stats.pairs_tested == 45


@pytest.mark.parametrize("M,N,n_case", [(1, 1, 1), (2, 15, 7), (3, 16, 0), (5, 17, 17), (33, 31, 15), (40, 33, 16),
                                         (70, 127, 64), (129, 129, 1), (64, 513, 256), (65, 1025, 1000)])
def test_ragged_and_degenerate_shapes(orc, M, N, n_case):
    """Tables of one SNP or one sample, sizes around the 16-/32-/128-sample block edges, one empty class: layout, both scan
    kernels, every table overload, the screen and the G-test against the oracle."""
    codes, _ = orc.simulate(1000 + 7 * M + N, M, N, max(n_case, 1) if n_case < N else N, missing_rate=0.05)
    pheno = np.zeros(N, np.uint8)
    pheno[np.random.default_rng(N).permutation(N)[:n_case]] = 1
    rows = orc.pack_codes(codes)
    sel, nca, nco = orc.select(rows, N, pheno)
    assert (nca, nco) == (n_case, N - n_case)
    want = orc.cc_counts_selected(sel, nca, nco)
    with gw.GenoStore(M, N) as st:
        st.put_rows(rows)
        assert np.array_equal(st.get_rows(), rows)
        st.select_case_control(pheno)
        first = st.marginal_scan()                                  # masked scan on the raw rows
        assert np.array_equal(st.get_selected_rows(), sel)          # K0
        second = st.marginal_scan()                                 # compacted rows
        for o in (first, second):
            assert np.array_equal(o["counts"], want)
        assert first["mi"].tobytes() == second["mi"].tobytes() and first["stats"].tobytes() == second["stats"].tobytes()
        mar = orc.margins(sel, nca, nco)
        assert np.array_equal(first["mi"]["dPbc"], mar["pbc"]) and np.array_equal(first["mi"]["dPca"], mar["pca"])
        assert np.array_equal(st.counts(1), orc.cc_counts_masked(rows, N, pheno))
        assert np.array_equal(st.counts(0)[:, :3], orc.counts_whole(rows, N)[:, :3])
        if M >= 2:
            pi, pj = np.triu_indices(M, 1)
            pi, pj = pi[:200].astype(np.uint32), pj[:200].astype(np.uint32)
            for mode in (0, 1, 2, 3):
                got = st.pair_tables(pi, pj, mode)
                for q in range(0, len(pi), 17):
                    ca, co = orc.pair_table(mode, int(pi[q]), int(pj[q]), rows=rows, sel=sel, n_samples=N, pheno=pheno,
                                            nca=nca, nco=nco, mar=mar)
                    assert np.array_equal(got[q, :16], ca) and np.array_equal(got[q, 16:], co), (mode, q)
            hits, stats = st.pairwise_scan(30.0)
            hi, hj, hs, _ = orc.boost_screen(sel, mar, nca, nco, 30.0)
            assert stats.pairs_tested == M * (M - 1) // 2
            assert np.array_equal(hits["i"], hi) and np.array_equal(hits["j"], hj)
            assert rel_close(hits["stat"], hs, REL_F64)


# ------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties (the oracle cannot run these in seconds)
# ------------------------------------------------------------------------------------------------------
def test_full_size_marginal_scan_properties(orc):
    """configs[1]: 5 000 cases / 5 000 controls x 500 000 SNPs, generated in HBM."""
    M, N, NCASE, SEED = 500_000, 10_000, 5_000, 20121127
    with gw.GenoStore(M, N) as st:
        st.simulate(SEED)
        pheno = gw.simulate_phenotype(SEED, N, NCASE)
        st.select_case_control(pheno)
        out = st.marginal_scan(mi=False)
        c = out["counts"].astype(np.int64)
        assert np.all(c[:, :4].sum(1) == NCASE) and np.all(c[:, 4:].sum(1) == N - NCASE)
        assert np.all(c[:, 3] == 0) and np.all(c[:, 7] == 0)                     # the generator draws no missing calls
        # allele-count conservation: the generator placed exactly floor(p * 2N) minor alleles in each SNP
        first = c[:, 0] + c[:, 4]                                                # first-seen homozygote
        het = c[:, 1] + c[:, 5]
        second = c[:, 2] + c[:, 6]
        minor = np.minimum(2 * first + het, 2 * second + het)
        idx = np.arange(0, M, 997)
        want = np.array([orc.minor_alleles(SEED, 1, N, first_snp=int(s))[0] for s in idx], np.int64)
        want = np.minimum(want, 2 * N - want)                                    # bins reach 50.999 %: the drawn allele can be the major one
        assert np.array_equal(minor[idx], want)
        # spot rows against the oracle end to end (pack -> select -> count -> statistics)
        rows = st.get_rows(123_456, 4)
        sel, nca, nco = orc.select(rows, N, pheno)
        assert np.array_equal(orc.cc_counts_selected(sel, nca, nco), out["counts"][123_456:123_460])
        s = out["stats"][123_456]
        x, p = orc.chi2_allelic(out["counts"][123_456, :4], out["counts"][123_456, 4:])
        assert rel_close(s["chi2_allelic"], x, REL_F64)
        # idempotence and device-output path
        again = st.marginal_scan(mi=False)
        assert np.array_equal(again["counts"], out["counts"]) and again["stats"].tobytes() == out["stats"].tobytes()


def test_full_size_pairwise_properties(orc):
    """configs[2]: 2 000 / 2 000 samples x 50 000 SNPs = 1 249 975 000 pairs."""
    M, N, NCASE, SEED = 50_000, 4_000, 2_000, 20121127
    with gw.GenoStore(M, N) as st:
        st.simulate(SEED)
        pheno = gw.simulate_phenotype(SEED, N, NCASE)
        st.select_case_control(pheno)
        hits, stats = st.pairwise_scan(30.0)
        assert stats.pairs_tested == M * (M - 1) // 2 == 1_249_975_000
        assert stats.word_cells == stats.pairs_tested * 4 * (63 + 63)
        key = hits["i"].astype(np.int64) * M + hits["j"]
        assert np.all(np.diff(key) > 0) and np.all(hits["i"] < hits["j"]) and np.all(hits["stat"] > 30.0)
        # every reported statistic is reproduced by the stand-alone fp64 kernel
        assert np.array_equal(st.ksa(hits["i"], hits["j"]), hits["stat"])
        # brute force with the oracle on a corner of the pair space: identical hit set and statistics
        B = 1500
        sel = st.get_selected_rows(0, B)
        mar = orc.margins(sel, NCASE, N - NCASE)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, NCASE, N - NCASE, 30.0)
        corner = hits[(hits["i"] < B) & (hits["j"] < B)]
        assert np.array_equal(corner["i"], hi) and np.array_equal(corner["j"], hj) and rel_close(corner["stat"], hs, REL_F64)
        # two shards partition the result; a second run is identical
        a, sa = st.pairwise_scan(30.0, shard=0, n_shards=2)
        b, sb = st.pairwise_scan(30.0, shard=1, n_shards=2)
        assert sa.pairs_tested + sb.pairs_tested == stats.pairs_tested
        assert np.array_equal(np.sort(np.concatenate([a, b]), order=["i", "j"]), hits)
