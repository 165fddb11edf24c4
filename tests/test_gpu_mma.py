"""GPU parity tests of the tensor-core pair-screen engine (libgwaspp_b200/csrc/pairwise_mma.cu).

The engine counts the four corner cells of getCaseControlContingencyTable's no-missing shortcut
(compressed_genotype_table5.cpp:1069-1144) with tcgen05 kind::i8 MMAs. Bars: counts bit-exact against the
per-call pair tables (themselves pinned to the oracle and the reference's golden vectors); hit sets and
fp64 statistics identical to the AND+POPC engine and to the oracle.
"""
import numpy as np
import pytest

import libgwaspp_b200 as gw
from helpers import planted_cohort, rel_close

pytestmark = pytest.mark.gpu


def make_store(orc, codes, pheno):
    M, N = codes.shape
    st = gw.GenoStore(M, N)
    st.put_rows(orc.pack_codes(codes))
    st.select_case_control(pheno)
    return st


def expected_tile(st, M, I, J):
    """[64, 128, 2, 4] corner counts from the per-call margins-overload tables (zero outside the table)."""
    gi = np.arange(64 * I, 64 * I + 64)
    gj = np.arange(128 * J, 128 * J + 128)
    A, B = np.meshgrid(gi, gj, indexing="ij")
    ok = (A < M) & (B < M)
    t = st.pair_tables(A[ok], B[ok], mode=3).reshape(-1, 2, 16)
    exp = np.zeros((64, 128, 2, 4), np.uint32)
    exp[ok] = t[:, :, [0, 2, 8, 10]]          # AA_BB, AA_bb, aa_BB, aa_bb of the 4x4 table
    return exp


@pytest.mark.parametrize("seed,M,N,ncase", [
    (21, 300, 700, 333),        # ragged everything: 5 A-blocks, 3 B-blocks, class sizes not multiples of 32
    (22, 256, 4096, 2048),      # exact multiples: 32 sample blocks of 128 bytes, several passes over the stage ring
    (23, 70, 40, 9),            # tiny: one sample block per class, mostly padding
])
def test_mma_tile_counts_match_pair_tables(orc, seed, M, N, ncase):
    codes, pheno = orc.simulate(seed, M, N, ncase)
    with make_store(orc, codes, pheno) as st:
        TA, TB = (M + 63) // 64, (M + 127) // 128
        tiles = [(I, J) for J in range(TB) for I in range(TA) if I // 2 <= J]
        for I, J in tiles:
            got = st.mma_tile_counts(I, J)
            assert np.array_equal(got, expected_tile(st, M, I, J)), f"tile ({I}, {J})"


@pytest.mark.parametrize("seed,M,N,ncase,miss,planted", [
    (31, 900, 1500, 700, 0.0, 12),     # two bands of A-blocks, ragged edges
    (32, 2500, 500, 250, 0.0, 10),     # 40 A-blocks: three bands, partial last band
    (33, 333, 900, 371, 0.01, 8),      # mixed: tensor-core tiles + 9-cell AND+POPC tiles
])
def test_engines_agree_with_each_other_and_the_oracle(orc, seed, M, N, ncase, miss, planted):
    codes, pheno = planted_cohort(orc, seed, M, N, ncase, miss, planted)
    if miss > 0:
        codes[:192][codes[:192] == 3] = 0           # the first three 64-SNP blocks completely called
    with make_store(orc, codes, pheno) as st:
        st.set_pair_engine(1)
        h1, s1 = st.pairwise_scan(30.0)
        st.set_pair_engine(2)
        h2, s2 = st.pairwise_scan(30.0)
        assert (s1.engine, s2.engine) == (1, 2)
        assert s1.pairs_tested == s2.pairs_tested == M * (M - 1) // 2
        assert np.array_equal(h1, h2)               # (i, j, fp64 stat) records, bit for bit
        sel = st.get_selected_rows()
        mar = orc.margins(sel, st.n_case, st.n_ctrl)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, st.n_case, st.n_ctrl, 30.0)
        assert len(hi) >= planted // 2
        assert np.array_equal(h2["i"], hi) and np.array_equal(h2["j"], hj) and rel_close(h2["stat"], hs, 1e-12)
        # shards of the tensor-core schedule partition the pair space
        parts = [st.pairwise_scan(30.0, shard=k, n_shards=3) for k in range(3)]
        assert sum(p[1].pairs_tested for p in parts) == M * (M - 1) // 2
        assert np.array_equal(np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"]), h2)
        # low threshold: many candidates through the fp32 epilogue
        st.set_pair_engine(1)
        l1, _ = st.pairwise_scan(10.0)
        st.set_pair_engine(2)
        l2, _ = st.pairwise_scan(10.0)
        assert len(l2) > len(h2) and np.array_equal(l1, l2)
        st.set_pair_engine(0)
        # the fp32 epilogue stays inside its safety margin (0.5)
        if miss == 0:
            ii, jj = np.triu_indices(min(M, 200), 1)
            both, f64 = st.ksa_screen_mma_f32(ii, jj), st.ksa(ii, jj)
            f32, ub = both[:, 0], both[:, 1]
            ok = ~np.isnan(f64)
            assert np.array_equal(np.isnan(f32), np.isnan(f64)) and np.array_equal(np.isnan(ub), np.isnan(f64))
            assert np.max(np.abs(f32[ok] - f64[ok])) < 0.125
            # the pre-filter is an upper bound of the statistic (up to fp32 rounding, far inside the margin)
            assert np.min(ub[ok] - f64[ok]) > -0.05
            assert np.mean(ub[ok] > 29.5) < 0.05            # ... and a useful one


def test_engine_selection_errors(orc):
    codes, pheno = orc.simulate(5, 64, 200, 100)
    with make_store(orc, codes, pheno) as st:
        with pytest.raises(gw.GwasDevError):
            st.set_pair_engine(3)
        st.set_pair_engine(2)
        hits, stats = st.pairwise_scan(30.0)
        assert stats.engine == 2 and stats.pairs_tested == 64 * 63 // 2


def test_four_plane_engine_on_a_cohort_where_every_block_has_missing_calls(orc):
    """Real genotype data: every 64-SNP block has missing calls, so the whole screen is the reference's 9-cell branch
    (compressed_genotype_table5.cpp:1000-1067). The four-plane tensor-core kernel (planes aa, bb, xx) must return the
    AND+POPC kernel's records bit for bit and the oracle's hit set, also sharded."""
    M, N, ncase = 700, 1500, 640
    codes, pheno = planted_cohort(orc, 77, M, N, ncase, 0.02, 6)
    with make_store(orc, codes, pheno) as st:
        h4, s4 = st.pairwise_scan(30.0)                      # default engine: tensor cores
        assert s4.engine == 2 and s4.tiles_nine_cell == s4.tiles > 0 and s4.pairs_tested == M * (M - 1) // 2
        st.set_pair_engine(1)
        h1, s1 = st.pairwise_scan(30.0)
        assert s1.engine == 1 and np.array_equal(h1, h4)
        st.set_pair_engine(0)
        sel = st.get_selected_rows()
        mar = orc.margins(sel, st.n_case, st.n_ctrl)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, st.n_case, st.n_ctrl, 30.0)
        assert len(hi) >= 3 and np.array_equal(h4["i"], hi) and np.array_equal(h4["j"], hj) and rel_close(h4["stat"], hs, 1e-12)
        parts = [st.pairwise_scan(30.0, shard=k, n_shards=5) for k in range(5)]
        assert sum(p[1].pairs_tested for p in parts) == M * (M - 1) // 2
        assert np.array_equal(np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"]), h4)
        l4, _ = st.pairwise_scan(-1e9)                       # every pair with a finite statistic goes through epilogue and re-score
        st.set_pair_engine(1)
        l1, _ = st.pairwise_scan(-1e9)
        assert len(l4) > 100 * len(h4) and np.array_equal(l1, l4)


@pytest.mark.parametrize("M,N,ncase,miss,popc_ok", [(200, 40_000, 20_000, 0.0, True), (130, 150_000, 70_000, 0.0, False),
                                                     (200, 40_000, 20_000, 0.01, True), (130, 150_000, 70_000, 0.005, False)])
def test_split_class_planes_for_cohorts_beyond_the_packed_accumulator(orc, M, N, ncase, miss, popc_ok):
    """n_case >= 16384: the two-plane kernel's int32 accumulator (n_case + 2^14 n_ctrl) no longer fits. Complete cohorts: the
    four-plane kernel with one pair of planes per class counts both classes as full int32 products. Cohorts with missing
    calls: planes aa, bb, xx with cases and controls accumulated into separate TMEM accumulators. Records must equal the
    AND+POPC engine's (while that one still applies: classes below 65536) and the oracle's hit set."""
    codes, pheno = planted_cohort(orc, 91, M, N, ncase, miss, 5)
    with make_store(orc, codes, pheno) as st:
        h, s = st.pairwise_scan(30.0)
        assert s.engine == 2 and s.pairs_tested == M * (M - 1) // 2
        sel = st.get_selected_rows()
        mar = orc.margins(sel, st.n_case, st.n_ctrl)
        hi, hj, hs, _ = orc.boost_screen(sel, mar, st.n_case, st.n_ctrl, 30.0)
        assert len(hi) >= 3 and np.array_equal(h["i"], hi) and np.array_equal(h["j"], hj) and rel_close(h["stat"], hs, 1e-12)
        parts = [st.pairwise_scan(30.0, shard=k, n_shards=2) for k in range(2)]
        assert np.array_equal(np.sort(np.concatenate([p[0] for p in parts]), order=["i", "j"]), h)
        low, _ = st.pairwise_scan(-1e9)
        assert len(low) == M * (M - 1) // 2 or len(low) > 100 * len(h)
        st.set_pair_engine(1)
        if popc_ok:
            h1, s1 = st.pairwise_scan(30.0)
            l1, _ = st.pairwise_scan(-1e9)
            assert s1.engine == 1 and np.array_equal(h1, h) and np.array_equal(l1, low)
        else:
            with pytest.raises(gw.GwasDevError, match="65535"):
                st.pairwise_scan(30.0)
