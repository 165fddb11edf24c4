"""Sample-block streaming of the marginal scan (BASELINE configs[4]): counts are additive over disjoint sample
blocks of one cohort, so a cohort that is loaded (or generated) block by block gives bit-identical counts,
marginal_information and statistics to one scan over the whole table -- provided the genotype labels stay
first-seen over the WHOLE row (genetics/genotype/common_genotype.h:257-304), which is what the label state carried
from block to block is for.
"""
import numpy as np
import pytest

import libgwaspp_b200 as gw

pytestmark = pytest.mark.gpu
SEED = 20121127
TXT = {0: b"AA", 1: b"AC", 2: b"CC", 3: b"00"}


def whole_scan(M, N, ncase, miss):
    pheno = gw.simulate_phenotype(SEED, N, ncase)
    with gw.GenoStore(M, N) as st:
        st.simulate(SEED, missing_rate=miss)
        rows = st.get_rows()
        st.select_case_control(pheno)
        return pheno, rows, st.marginal_scan()


@pytest.mark.parametrize("M,N,ncase,miss,blocks", [
    (700, 5000, 2100, 0.0, [1024, 1024, 1024, 1024, 904]),
    (257, 3001, 1499, 0.02, [37, 1500, 1, 1463]),          # ragged blocks, missing calls, a one-sample block
])
def test_generated_blocks_accumulate_to_the_whole_scan(M, N, ncase, miss, blocks):
    assert sum(blocks) == N
    pheno, rows, ref = whole_scan(M, N, ncase, miss)
    acc = np.zeros((M, 8), np.uint32)
    s0 = 0
    for nb in blocks:
        with gw.GenoStore(M, nb) as st:
            st.simulate_block(SEED, s0, N, missing_rate=miss)
            # the block's raw planes are the whole table's columns [s0, s0 + nb), labels included
            blk = st.get_rows()
            P, Pb = gw.plane_blocks(N), gw.plane_blocks(nb)
            for plane in range(2):
                full = np.unpackbits(rows[:, 1 + plane * P:1 + (plane + 1) * P].view(np.uint8), axis=1, bitorder="little")[:, s0:s0 + nb]
                part = np.unpackbits(blk[:, 1 + plane * Pb:1 + (plane + 1) * Pb].view(np.uint8), axis=1, bitorder="little")[:, :nb]
                assert np.array_equal(full, part)
            st.select_case_control(pheno[s0:s0 + nb])
            st.marginal_accumulate(acc)
        s0 += nb
    assert np.array_equal(blk[:, 0], rows[:, 0])                     # header word after the last block = whole-row header
    assert np.array_equal(acc, ref["counts"])
    mi, stats = gw.marginal_finalize(acc)
    assert mi.tobytes() == ref["mi"].tobytes() and stats.tobytes() == ref["stats"].tobytes()


def test_text_blocks_keep_first_seen_labels(orc):
    M, N, ncase = 60, 400, 170
    codes, pheno = orc.simulate(7, M, N, ncase, missing_rate=0.05)
    with gw.GenoStore(M, N) as st:
        st.put_rows(orc.pack_codes(codes))
        st.select_case_control(pheno)
        ref = st.marginal_scan()
    acc = np.zeros((M, 8), np.uint32)
    state = [0] * M
    for s0, nb in ((0, 130), (130, 7), (137, 263)):
        rows = []
        for r in range(M):
            line = b"\t".join(TXT[int(c)] for c in codes[r, s0:s0 + nb])
            row, state[r] = gw.pack_row_text_block(line, nb, state[r])
            rows.append(row)
        with gw.GenoStore(M, nb) as st:
            st.put_rows(np.stack(rows))
            st.select_case_control(pheno[s0:s0 + nb])
            st.marginal_accumulate(acc)
    assert np.array_equal(acc, ref["counts"])
    mi, stats = gw.marginal_finalize(acc)
    assert mi.tobytes() == ref["mi"].tobytes() and stats.tobytes() == ref["stats"].tobytes()


def test_biobank_scale_resident_scan_properties():
    """configs[4] resident: 100 000 cases / 100 000 controls x 1 000 000 SNPs, 50 GB of 2-bit genotypes in HBM."""
    M, N, NCASE = 1_000_000, 200_000, 100_000
    with gw.GenoStore(M, N) as st:
        st.simulate(SEED)
        pheno = gw.simulate_phenotype(SEED, N, NCASE)
        st.select_case_control(pheno)
        out = st.marginal_scan(mi=False)
        c = out["counts"].astype(np.int64)
        assert np.all(c[:, :4].sum(1) == NCASE) and np.all(c[:, 4:].sum(1) == N - NCASE)
        assert np.all(c[:, 3] == 0) and np.all(c[:, 7] == 0)                      # the simulator has no missing calls
        # allele conservation against the generator: minor alleles = floor(p * 2N) by construction, so the
        # pooled minor-allele frequency equals floor(p 2N) / 2N with p on the 1e-5 grid of simulate_data.cpp:191-192
        het, hom1, hom2 = c[:, 1] + c[:, 5], c[:, 0] + c[:, 4], c[:, 2] + c[:, 6]
        minor = np.minimum(2 * hom1 + het, 2 * hom2 + het)
        assert np.allclose(out["stats"]["maf_pooled"], minor / (2.0 * N), rtol=0, atol=1e-15)
        assert minor.max() <= N and (minor > 0).mean() > 0.5
        # spot rows against an independent recount of the raw planes on the host
        for r in (0, 123_456, M - 1):
            row = st.get_rows(r, 1)[0]
            P = gw.plane_blocks(N)
            p1 = np.unpackbits(row[1:1 + P].view(np.uint8), bitorder="little")[:N].astype(bool)
            p2 = np.unpackbits(row[1 + P:1 + 2 * P].view(np.uint8), bitorder="little")[:N].astype(bool)
            ca = pheno.astype(bool)
            exp = [(p1 & ~p2 & ca).sum(), (p2 & ~p1 & ca).sum(), (p1 & p2 & ca).sum(), (~p1 & ~p2 & ca).sum(),
                   (p1 & ~p2 & ~ca).sum(), (p2 & ~p1 & ~ca).sum(), (p1 & p2 & ~ca).sum(), (~p1 & ~p2 & ~ca).sum()]
            assert list(out["counts"][r]) == exp
