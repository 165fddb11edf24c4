"""GPU parity tests added in round 2: the north star's target shape, whole-problem engine equality, the bound pre-filter's
slack at scale, bounded (top-k) results, compact marginal outputs, on-the-fly masks that leave the selection alone, rows
with set padding bits, and the multi-device driver (tests that need two GPUs skip on a one-GPU box).

Bars as everywhere: counts, tables, hit sets bit-exact; fp64 statistics within 1e-12 relative of the oracle.
Nothing here reads /root/reference.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import libgwaspp_b200 as gw
from helpers import planted_cohort, rel_close

pytestmark = pytest.mark.gpu
REL_F64 = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    return int(gw.load_library().gwasdev_device_count())


def make_store(orc, codes, pheno=None, device=0):
    M, N = codes.shape
    st = gw.GenoStore(M, N, device=device)
    st.put_rows(orc.pack_codes(codes))
    if pheno is not None:
        st.select_case_control(pheno)
    return st


def top_k_of(hits, k):
    """the k largest statistics, ties at the k-th place to the smaller (i, j); returned in (i, j) order"""
    order = np.lexsort((hits["j"], hits["i"], -hits["stat"]))[:k]
    return np.sort(hits[order], order=["i", "j"])


# ------------------------------------------------------------------------------------------------------
# the north star's target shape: 5 000 cases / 5 000 controls (80 sample blocks through the 3-stage ring)
# ------------------------------------------------------------------------------------------------------
def test_target_shape_5000_cases_5000_controls_against_the_oracle(orc):
    M, N, NCASE = 2000, 10_000, 5_000
    codes, pheno = planted_cohort(orc, 510, M, N, NCASE, 0.0, 12)
    with make_store(orc, codes, pheno) as st:
        st.set_pair_engine(2)
        h2, s2 = st.pairwise_scan(30.0)
        assert s2.engine == 2 and s2.pairs_tested == M * (M - 1) // 2 and s2.tiles_nine_cell == 0
        st.set_pair_engine(1)
        h1, s1 = st.pairwise_scan(30.0)
        assert s1.engine == 1 and np.array_equal(h1, h2)                      # record for record
        sel = st.get_selected_rows()
        mar = orc.margins(sel, st.n_case, st.n_ctrl)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, st.n_case, st.n_ctrl, 30.0)
        assert len(hi) >= 8
        assert np.array_equal(h2["i"], hi) and np.array_equal(h2["j"], hj) and rel_close(h2["stat"], hs, REL_F64)
        # raw corner counts of one tile against the per-call tables at this row length
        got = st.mma_tile_counts(3, 7)
        gi, gj = np.meshgrid(np.arange(192, 256), np.arange(896, 1024), indexing="ij")
        t = st.pair_tables(gi.ravel(), gj.ravel(), mode=3).reshape(64, 128, 2, 16)
        assert np.array_equal(got, t[:, :, :, [0, 2, 8, 10]])
        # the same cohort with missing calls: four-plane kernel at 80 sample blocks
    codes[:, ::7][codes[:, ::7] == 1] = 3
    with make_store(orc, codes, pheno) as st:
        h4, s4 = st.pairwise_scan(30.0)
        assert s4.engine == 2 and s4.tiles_nine_cell > 0
        st.set_pair_engine(1)
        h9, _ = st.pairwise_scan(30.0)
        assert np.array_equal(h4, h9)
        sel = st.get_selected_rows()
        mar = orc.margins(sel, st.n_case, st.n_ctrl)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, st.n_case, st.n_ctrl, 30.0)
        assert np.array_equal(h4["i"], hi) and np.array_equal(h4["j"], hj) and rel_close(h4["stat"], hs, REL_F64)


def test_full_configs2_both_engines_record_for_record():
    """configs[2], all 1 249 975 000 pairs: the tensor-core engine (bound pre-filter, fp32 exact pass) and the AND+POPC
    engine (no pre-filter) must report the same records -- a false negative of the bound would show up here, wherever it
    sits in the pair space. Also with a sprinkle of missing calls (mixed two-plane / four-plane tiles) and sharded."""
    M, N, NC, SEED = 50_000, 4_000, 2_000, 20121127
    for missing in (0.0, 0.0002):
        with gw.GenoStore(M, N) as st:
            st.simulate(SEED, missing_rate=missing)
            st.select_case_control(gw.simulate_phenotype(SEED, N, NC))
            whole, s = st.pairwise_scan(30.0)
            assert s.engine == 2 and s.pairs_tested == M * (M - 1) // 2
            st.set_pair_engine(1)
            popc, s1 = st.pairwise_scan(30.0)
            assert s1.engine == 1 and len(whole) > 1000
            assert np.array_equal(whole, popc)
            st.set_pair_engine(0)
            parts = [st.pairwise_scan(30.0, shard=k, n_shards=8) for k in range(8)]
            assert sum(p[1].pairs_tested for p in parts) == M * (M - 1) // 2
            assert np.array_equal(np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"]), whole)
            # bounded result on the same problem: the 1 000 strongest pairs
            top, _ = st.pairwise_topk(1000, 30.0)
            assert np.array_equal(top, top_k_of(whole, 1000))


@pytest.mark.parametrize("N,ncase,M", [(4_000, 2_000, 4500), (10_000, 5_000, 4500), (146_000, 16_000, 4500)])
def test_bound_prefilter_slack_over_many_pairs(orc, N, ncase, M):
    """The cheap upper bound of the tensor-core epilogue (rcp.approx / lg2.approx, fp32 per-SNP constants of magnitude
    N log2 N) against the fp64 statistic over >= 1e7 random pairs (all pairs of M SNPs where that is fewer): it may fall
    below the statistic by fp32 rounding only, far inside the screening margin max(0.5, 1e-4 N) the kernel subtracts
    from the threshold -- asserted at half the margin."""
    rng = np.random.default_rng(N)
    with gw.GenoStore(M, N) as st:
        st.simulate(20121127 + N)
        st.select_case_control(gw.simulate_phenotype(7, N, ncase))
        n_pairs = min(10_000_000, M * (M - 1) // 2)
        if n_pairs == M * (M - 1) // 2:
            pi, pj = np.triu_indices(M, 1)
        else:
            pi = rng.integers(0, M - 1, n_pairs)
            pj = pi + 1 + (rng.integers(0, 1 << 30, n_pairs) % (M - 1 - pi))
        margin = max(0.5, 1e-4 * N)
        worst_ub, worst_f32 = 0.0, 0.0
        for b in range(0, n_pairs, 2_000_000):
            a, c = pi[b:b + 2_000_000].astype(np.uint32), pj[b:b + 2_000_000].astype(np.uint32)
            both, f64 = st.ksa_screen_mma_f32(a, c), st.ksa(a, c)
            f32, ub = both[:, 0].astype(np.float64), both[:, 1].astype(np.float64)
            ok = ~np.isnan(f64)
            assert np.array_equal(np.isnan(ub), np.isnan(f64))
            worst_ub = min(worst_ub, float(np.min(ub[ok] - f64[ok])))
            worst_f32 = max(worst_f32, float(np.max(np.abs(f32[ok] - f64[ok]))))
        assert worst_ub >= -margin / 2, (worst_ub, margin)
        assert worst_f32 <= margin / 2, (worst_f32, margin)
        print(f"N={N}: {n_pairs} pairs, min(ub - f64) = {worst_ub:.4f}, max|f32 - f64| = {worst_f32:.4f}, margin {margin}")


# ------------------------------------------------------------------------------------------------------
# bounded results: top-k with a rising device-wide threshold, candidate-buffer overflow paths
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("miss,engine", [(0.0, 0), (0.0, 1), (0.02, 0), (0.02, 1)])
def test_topk_equals_the_k_largest_of_the_full_list(orc, miss, engine):
    M, N, NCASE = 640, 900, 420
    codes, pheno = planted_cohort(orc, 801, M, N, NCASE, miss, 6)
    # a threshold almost every pair passes. With missing calls the reference's statistic (per-SNP margins against a table that
    # lacks the incomplete samples, epistasis_func.cpp:424-470) is far below zero for ordinary pairs, so "almost every pair"
    # needs a threshold far below zero too -- which also takes the histogram through the negative half of fp32.
    thr = 2.0 if miss == 0 else -1e9
    with make_store(orc, codes, pheno) as st:
        st.set_pair_engine(engine)
        full, _ = st.pairwise_scan(thr, capacity=M * M)
        assert len(full) > 0.5 * M * (M - 1) // 2
        for k in (1, 7, 1000, len(full) + 5):
            top, s = st.pairwise_topk(k, thr)
            assert np.array_equal(top, top_k_of(full, k)), k
        # a candidate buffer far smaller than the list: the device-wide threshold has to rise while the kernel runs
        st.set_option(gw.OPT_CAND_CAPACITY, 4096)
        for k in (5, 300):
            top, s = st.pairwise_topk(k, thr)
            assert np.array_equal(top, top_k_of(full, k)), k
            assert s.candidates < len(full)                                   # pairs below the risen threshold were never appended
        # threshold mode with the same tiny buffer: the screen re-runs with the exact size and returns the same list
        again, s = st.pairwise_scan(thr, capacity=M * M)
        assert np.array_equal(again, full)
        st.set_option(gw.OPT_CAND_CAPACITY, 0)
        # shards: the top-k of the union of the shards' top-k lists is the global top-k
        parts = [st.pairwise_topk(50, thr, shard=r, n_shards=3)[0] for r in range(3)]
        assert np.array_equal(top_k_of(np.concatenate(parts), 50), top_k_of(full, 50))


# ------------------------------------------------------------------------------------------------------
# marginal scan: compact outputs, on-the-fly masks, padding bits
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name_seed,N,ncase", [(11, 1000, 480), (12, 2600, 1300)])
def test_compact_scan_outputs(orc, name_seed, N, ncase):
    M = 700
    codes, pheno = planted_cohort(orc, name_seed, M, N, ncase, 0.01, 4)
    idx = np.arange(N)                                                        # a strongly associated SNP: p far below 1e-30
    codes[5, (pheno == 1) & (idx % 10 < 3)] = 0
    codes[5, (pheno == 0) & (idx % 10 < 3)] = 2
    with make_store(orc, codes, pheno) as st:
        full = st.marginal_scan(mi=False)
        for first in (True, False):                                           # masked first scan, then whatever follows
            if first:
                st.select_case_control(pheno)
            rec, sig = st.marginal_scan_compact(p_threshold=1e-4)
            assert np.array_equal(rec["cases"], full["counts"][:, :4]) and np.array_equal(rec["controls"], full["counts"][:, 4:])
            for f in ("chi2_allelic", "p_allelic", "chi2_genotypic", "p_genotypic"):
                want = full["stats"][f].astype(np.float32)                    # the fp64 value rounded once
                assert np.array_equal(rec[f], want, equal_nan=True), f
            want_sig = np.flatnonzero((full["stats"]["p_allelic"] < 1e-4) | (full["stats"]["p_genotypic"] < 1e-4))
            assert 5 in want_sig and np.array_equal(sig["snp"], want_sig)
            for f in ("maf_pooled", "chi2_allelic", "p_allelic", "chi2_genotypic", "p_genotypic"):
                assert np.array_equal(sig[f], full["stats"][f][want_sig])      # the same fp64 bits
            assert np.array_equal(sig["df_genotypic"], full["stats"]["df_genotypic"][want_sig].astype(np.uint32))
            assert 0.0 < sig["p_allelic"][list(want_sig).index(5)] < 1e-30 and rec["p_allelic"][5] < 1e-30
        # sub-range, significant list only, and the overflow report
        rec, sig = st.marginal_scan_compact(100, 400, records=False, p_threshold=1e-4)
        assert rec is None and np.array_equal(sig["snp"], want_sig[(want_sig >= 100) & (want_sig < 400)])
        n = C.c_uint64()
        buf = np.zeros(1, gw.SIG_DTYPE)
        rc = st.L.gwasdev_marginal_scan_compact(st.h, 0, M, None, 1e-4, buf.ctypes.data_as(C.c_void_p), 1, C.byref(n), 0)
        assert rc == 4 and n.value == len(want_sig)                           # GWASDEV_EOVERFLOW with the needed size


def test_on_the_fly_masks_leave_the_selection_alone(orc):
    """select with set A, probe on the fly with set B (getCaseControlGenotypeDistribution(r, ccs, ..) :609-657,
    getCaseControlContingencyTable(i, j, ccs, ..) :806-895), go on using the pre-selected overloads: they must still see A
    (the reference's mask overloads never touch m_cases_controls), and nothing is re-compacted."""
    M, N = 200, 1000
    codes, pheno_a = orc.simulate(61, M, N, 450, missing_rate=0.02)
    rng = np.random.default_rng(3)
    pheno_b = rng.choice([0, 1, 2], size=N, p=[0.5, 0.3, 0.2]).astype(np.uint8)
    rows = orc.pack_codes(codes)
    with make_store(orc, codes, pheno_a) as st:
        sel_a = st.get_selected_rows()
        counts_a = st.counts(2)
        scan_a = st.marginal_scan()
        hits_a, _ = st.pairwise_scan(20.0)
        launches = gw.launch_count()
        st.set_stream_masks(pheno_b)
        assert np.array_equal(st.counts(1), orc.cc_counts_masked(rows, N, pheno_b))
        pi, pj = np.triu_indices(40, 1)
        t1 = st.pair_tables(pi, pj, mode=1)
        for q in range(0, len(pi), 97):
            ca, co = orc.pair_table(1, int(pi[q]), int(pj[q]), rows=rows, n_samples=N, pheno=pheno_b)
            assert np.array_equal(t1[q, :16], ca) and np.array_equal(t1[q, 16:], co)
        # the selection is still A's
        assert (st.n_case, st.n_ctrl) == (int((pheno_a == 1).sum()), int((pheno_a == 0).sum()))
        assert np.array_equal(st.counts(2), counts_a) and np.array_equal(st.get_selected_rows(), sel_a)
        again = st.marginal_scan()
        assert all(again[k].tobytes() == scan_a[k].tobytes() for k in again)
        h2, _ = st.pairwise_scan(20.0)
        assert np.array_equal(h2, hits_a)
        t3 = st.pair_tables(pi, pj, mode=3)
        sel, nca, nco = orc.select(rows, N, pheno_a)
        mar = orc.margins(sel, nca, nco)
        ca, co = orc.pair_table(3, 3, 17, sel=sel, nca=nca, nco=nco, mar=mar)
        q = int(np.flatnonzero((pi == 3) & (pj == 17))[0])
        assert np.array_equal(t3[q, :16], ca) and np.array_equal(t3[q, 16:], co)
        del launches
        # overlapping on-the-fly masks through the C-ABI: control xx uses the control count as given (:649-653)
        both = pheno_b.copy()
        ca_m, co_m = gw.stream_masks(both)
        for c in (3, 99, 500):
            ca_m[c >> 4] |= np.uint16(1 << (c & 15))
            co_m[c >> 4] |= np.uint16(1 << (c & 15))
        st.set_stream_masks(case_mask=ca_m, ctrl_mask=co_m)
        got = st.counts(1).astype(np.int64)
        bits = lambda m: np.unpackbits(m.view(np.uint8), bitorder="little")[:N].astype(bool)   # noqa: E731
        mca, mco = bits(ca_m), bits(co_m)
        for r in (0, 57, 199):
            for k, m in ((0, mca), (4, mco)):
                g = codes[r][m]
                assert list(got[r, k:k + 4]) == [int((g == 0).sum()), int((g == 1).sum()), int((g == 2).sum()), int((g == 3).sum())] or \
                    list(got[r, k:k + 4]) == [int((g == 2).sum()), int((g == 1).sum()), int((g == 0).sum()), int((g == 3).sum())]


def test_rows_with_set_padding_bits(orc):
    """rows handed to gwasdev_put_rows with garbage in the bits beyond sample N (the reference's row padding)
    must count like clean rows in every scan of a selection, masked or compacted."""
    M, N, NCASE = 150, 1003, 500                       # 1003 samples: 5 padding bits in the last 16-bit block, 3 whole padding blocks
    codes, pheno = orc.simulate(71, M, N, NCASE, missing_rate=0.01)
    rows = orc.pack_codes(codes)
    P = (rows.shape[1] - 1) // 2
    dirty = rows.copy()
    for pl in range(2):
        dirty[:, 1 + pl * P + N // 16] |= np.uint16(0xFFFF << (N % 16) & 0xFFFF)
        dirty[:, 1 + pl * P + N // 16 + 1: 1 + (pl + 1) * P] = 0xFFFF
    with gw.GenoStore(M, N) as clean, gw.GenoStore(M, N) as st:
        clean.put_rows(rows)
        clean.select_case_control(pheno)
        want = clean.marginal_scan()
        st.put_rows(dirty)
        assert np.array_equal(st.get_rows(), rows)                            # stored rows carry no padding bits
        st.select_case_control(pheno)
        for _ in range(3):                                                    # first scan of the table, cached totals, ...
            got = st.marginal_scan()
            assert all(got[k].tobytes() == want[k].tobytes() for k in want)
        assert np.array_equal(st.counts(0), clean.counts(0))
        st.set_select_mode(True)
        st.select_case_control(pheno)
        got = st.marginal_scan()
        assert all(got[k].tobytes() == want[k].tobytes() for k in want)


def test_reselection_scans_use_cached_row_totals(orc):
    """Partitioned cohorts: the first masked scan of a table writes per-SNP totals, later selections read them (three masked
    popcount streams). Counts must equal the oracle's for every selection, full range and sub-ranges, and after the rows change."""
    M, N = 333, 2100
    codes, _ = orc.simulate(81, M, N, 1000, missing_rate=0.02)
    rows = orc.pack_codes(codes)
    rng = np.random.default_rng(5)
    with gw.GenoStore(M, N) as st:
        st.put_rows(rows)
        for trial in range(4):
            pheno = (rng.random(N) < (0.2 + 0.2 * trial)).astype(np.uint8)
            st.select_case_control(pheno)
            sel, nca, nco = orc.select(rows, N, pheno)
            want = orc.cc_counts_selected(sel, nca, nco)
            if trial == 0:
                assert np.array_equal(st.marginal_scan(40, 300, mi=False, stats=False)["counts"], want[40:300])   # sub-range first: no totals yet
            assert np.array_equal(st.marginal_scan(mi=False, stats=False)["counts"], want)
            assert np.array_equal(st.marginal_scan(7, 111, mi=False, stats=False)["counts"], want[7:111])
        codes2, _ = orc.simulate(82, M, N, 1000, missing_rate=0.02)
        rows2 = orc.pack_codes(codes2)
        st.put_rows(rows2)                                                    # totals of the old table must not survive
        st.select_case_control(pheno)
        sel, nca, nco = orc.select(rows2, N, pheno)
        assert np.array_equal(st.marginal_scan(mi=False, stats=False)["counts"], orc.cc_counts_selected(sel, nca, nco))
        st.set_option(gw.OPT_ROW_TOTALS, 1)                                   # and the variant that never caches
        st.select_case_control(pheno)
        assert np.array_equal(st.marginal_scan(mi=False, stats=False)["counts"], orc.cc_counts_selected(sel, nca, nco))


@pytest.mark.parametrize("miss", [0.0, 0.02])
def test_hot_paths_never_compact(orc, miss):
    """Marginal scans of a partitioned cohort, the tensor-core screen (operands, fp64 re-score) and the G-test read the raw rows
    through the class masks: K0 runs only for the layout probes, and running it afterwards changes nothing."""
    M, N, NCASE = 500, 1100, 520
    codes, pheno = planted_cohort(orc, 313, M, N, NCASE, miss, 6)
    with make_store(orc, codes, pheno) as st, make_store(orc, codes) as eager:
        eager.set_select_mode(True)
        eager.select_case_control(pheno)
        assert eager.is_compacted() and not st.is_compacted()
        a, b = st.marginal_scan(), eager.marginal_scan()
        assert all(a[k].tobytes() == b[k].tobytes() for k in a)
        st.marginal_scan()                                                  # a second and third scan: still through the masks
        h, s = st.pairwise_scan(30.0)
        he, _ = eager.pairwise_scan(30.0)
        assert s.engine == 2 and len(h) >= 3 and np.array_equal(h, he)
        g, ge = st.gtest(h["i"], h["j"]), eager.gtest(he["i"], he["j"])
        assert np.array_equal(g[0], ge[0]) and np.array_equal(g[1], ge[1], equal_nan=True)
        pi, pj = np.triu_indices(60, 1)
        assert np.array_equal(st.ksa(pi, pj), eager.ksa(pi, pj), equal_nan=True)
        assert np.array_equal(st.pair_tables(pi, pj, mode=3), eager.pair_tables(pi, pj, mode=3))
        top, _ = st.pairwise_topk(3, 30.0)
        assert not st.is_compacted()                                        # none of the above needed the compacted layout
        sel = st.get_selected_rows()                                        # the layout probe does
        assert st.is_compacted() and np.array_equal(sel, eager.get_selected_rows())
        h2, _ = st.pairwise_scan(30.0)
        assert np.array_equal(h2, h)


@pytest.mark.parametrize("miss,shards", [(0.0, 1), (0.01, 1), (0.0, 5), (0.01, 3)])
def test_tile_feed_hands_out_every_tile_exactly_once(orc, miss, shards):
    """The CTA pairs of the tensor-core kernels draw their tiles from a device counter, so which pair computes which tile differs
    from launch to launch: twelve passes over a cohort of 40 x 40 tile blocks (820 tiles of 128 SNPs, more than the 74 pairs;
    with missing calls the four-plane kernel's 64-SNP tiles as well) must give the same pair count, candidates and records
    every time, shard by shard, and the shards together the whole list."""
    # complete data: threshold 25 over 13 M pairs; with missing calls the statistic of ordinary pairs is far below zero, so every
    # pair with a statistic of a smaller cohort (136 four-plane tiles) is kept: the list then checks the coverage pair by pair
    M, N, NCASE, thr = (5100, 1200, 590, 25.0) if miss == 0.0 else (1000, 1200, 590, -1e9)
    codes, pheno = planted_cohort(orc, 777, M, N, NCASE, miss, 10)
    with make_store(orc, codes, pheno) as st:
        st.set_pair_engine(2)
        whole, s0 = st.pairwise_scan(thr)
        assert s0.engine == 2 and s0.pairs_tested == M * (M - 1) // 2
        assert len(whole) >= (10 if miss == 0.0 else 100_000)      # (pairs with an empty genotype column have no statistic: NaN, never kept)
        for rep in range(12):
            parts, pairs = [], 0
            for sh in range(shards):
                h, s = st.pairwise_scan(thr, shard=sh, n_shards=shards)
                parts.append(h); pairs += s.pairs_tested
            got = np.sort(np.concatenate(parts), order=["i", "j"])
            assert pairs == s0.pairs_tested and np.array_equal(got, whole), rep
        st.set_pair_engine(1)
        h1, _ = st.pairwise_scan(thr)
        assert np.array_equal(h1, whole)


def test_tensor_peak_probe_is_plausible():
    burst, sustained = gw.i8_peak(0)
    assert 500.0 < sustained <= burst * 1.02 and burst < 5000.0, (burst, sustained)   # nominal dense int8: 4 500 TOP/s


def test_l2_delivery_probe_is_plausible():
    """What the L2 hands to the SMs (the rate the screens' operand streams run at) lies between the HBM read rate and the
    SMs' own load rate (148 x 128 B/clk); a buffer larger than the L2 falls back to the HBM rate and is refused."""
    l2, hbm = gw.l2_read_peak(0, 48 << 20, 100), gw.hbm_read_peak(0, 1 << 30)
    assert hbm < l2 < 37_000.0 and 3_000.0 < hbm < 8_500.0, (l2, hbm)
    with pytest.raises(gw.GwasDevError):
        gw.l2_read_peak(0, 256 << 20, 10)


# ------------------------------------------------------------------------------------------------------
# several devices in one process
# ------------------------------------------------------------------------------------------------------
needs_two = pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs in this process")


@needs_two
@pytest.mark.parametrize("miss", [0.0, 0.01])
def test_multi_device_driver_equals_one_device(orc, miss):
    M, N, NCASE = 1500, 1200, 560
    codes, pheno = planted_cohort(orc, 901, M, N, NCASE, miss, 10)
    nd = min(n_gpus(), 4)
    with make_store(orc, codes, pheno) as st:
        one, _ = st.pairwise_scan(25.0)
        assert len(one) >= 5
        others = [st.replicate(d) for d in range(1, nd)]
        try:
            for o in others:
                assert np.array_equal(o.get_rows(), st.get_rows()) and (o.n_case, o.n_ctrl) == (st.n_case, st.n_ctrl)
                assert np.array_equal(o.get_selected_rows(0, 50), st.get_selected_rows(0, 50))
            stores = [st] + others
            for gather in ("nccl", "peer"):
                many, stats = gw.pairwise_scan_multi(stores, 25.0, gather=gather)
                assert np.array_equal(many, one), gather
                assert sum(s.pairs_tested for s in stats) == M * (M - 1) // 2
            lo = 8.0 if miss == 0 else -1e9              # with missing calls ordinary pairs score far below zero (see the top-k test)
            low, _ = st.pairwise_scan(lo, capacity=M * M)
            many, _ = gw.pairwise_scan_multi(stores, lo, capacity=M * M)
            assert len(low) > 50 * len(one) and np.array_equal(many, low)
            top, _ = gw.pairwise_scan_multi(stores, lo, top_k=200)
            assert np.array_equal(top, top_k_of(low, 200))
            s1, z1 = st.gtest(one["i"], one["j"])
            s2, z2 = gw.gtest_multi(stores, one["i"], one["j"])
            assert np.array_equal(s1, s2) and np.array_equal(z1, z2, equal_nan=True)
            # a new selection on every device, then again
            ph2 = 1 - pheno
            for s in stores:
                s.select_case_control(ph2)
            one2, _ = st.pairwise_scan(25.0)
            many2, _ = gw.pairwise_scan_multi(stores, 25.0)
            assert np.array_equal(many2, one2)
        finally:
            for o in others:
                o.close()


@needs_two
def test_harness_on_two_devices_prints_the_one_device_output(orc, tmp_path):
    """gwas_b200 --test-boost-epi --devices 2: the C++ mirror's computeBoost through the in-library multi-device driver
    prints what one device prints (the reference's --test-boost-epi format), apart from the elapsed-time lines."""
    M, N, NCASE = 700, 800, 380
    codes, pheno = planted_cohort(orc, 77, M, N, NCASE, 0.0, 8)
    txt = {0: "A A", 1: "A C", 2: "C C", 3: "0 0"}
    tped, tfam = tmp_path / "c.tped", tmp_path / "c.tfam"
    with open(tped, "w") as f:
        for r in range(M):
            f.write(f"1 rs{r} 0 {r} " + " ".join(txt[int(c)] for c in codes[r]) + "\n")
    with open(tfam, "w") as f:
        for i, ph in enumerate(pheno):
            f.write(f"F{i} I{i} 0 0 1 {int(ph)}\n")
    exe = os.path.join(ROOT, "libgwaspp_b200", "gwas_b200")
    outs = []
    for nd in (1, 2):
        out = tmp_path / f"o{nd}.txt"
        subprocess.check_call([exe, "-g", str(tped), "-p", str(tfam), "-o", str(out), "--devices", str(nd), "--test-boost-epi"],
                              stdout=subprocess.DEVNULL)
        lines = [ln for ln in out.read_text().splitlines() if not re.fullmatch(r"\d+\.\d{9}s", ln.strip())]   # drop the lapse line
        outs.append(lines)
    assert outs[0] == outs[1] and sum("\t" in ln for ln in outs[0]) >= 4


def test_harness_comp_level_is_explicit(tmp_path):
    """--comp-level selects a HOST layout in the reference; the harness accepts the bit-plane levels (3, 4, 5), refuses the
    others, and refuses what the reference's level-3 table cannot run (its margins overloads assert(false))."""
    exe = os.path.join(ROOT, "libgwaspp_b200", "gwas_b200")
    tped, tfam = tmp_path / "t.tped", tmp_path / "t.tfam"
    tped.write_text("1 rs0 0 1 A A A C C C A A\n1 rs1 0 2 A C A A C C A C\n")
    tfam.write_text("F0 I0 0 0 1 1\nF1 I1 0 0 1 0\nF2 I2 0 0 1 1\nF3 I3 0 0 1 0\n")
    base = [exe, "-g", str(tped), "-p", str(tfam)]
    assert subprocess.run(base + ["--comp-level", "5", "--test-inline-maf"], capture_output=True).returncode == 0
    assert subprocess.run(base + ["--comp-level", "3", "--test-inline-maf"], capture_output=True).returncode == 0
    r = subprocess.run(base + ["--comp-level", "2", "--test-inline-maf"], capture_output=True, text=True)
    assert r.returncode != 0 and "comp-level" in r.stderr
    r = subprocess.run(base + ["--comp-level", "3", "--test-boost-epi"], capture_output=True, text=True)
    assert r.returncode != 0 and "margins" in r.stderr
