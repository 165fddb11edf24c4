"""GPU parity tests of the device-side loaders: TPED text, .gz and PLINK .bed parsed, labelled and
packed on the GPU must give the rows the reference's host loader gives -- byte for byte, headers included.

Checkers: tests/golden raw_rows (produced by the unmodified reference), the reference's own perl fixtures, and the
pinned oracle packer (oracle/gwas_oracle.c go_pack_row_text / go_pack_row_codes).
"""
import gzip
import os

import numpy as np
import pytest

import libgwaspp_b200 as gw
from helpers import load_golden

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CODE_TXT = np.array([b"A\tA\t", b"A\tC\t", b"C\tC\t", b"0\t0\t"])


def tped_bytes(codes, sep=b"\t", ids=None, eol=b"\n", spell=None):
    """TPED text of a code matrix (0 AA, 1 AC, 2 CC, 3 missing), built with numpy so that big tables stay cheap."""
    spell = spell or [b"A A", b"A C", b"C C", b"0 0"]
    lut = np.array([np.frombuffer(s.replace(b" ", sep) + sep, np.uint8).view(np.uint32)[0] for s in spell], np.uint32)
    out = []
    for r in range(codes.shape[0]):
        body = lut[codes[r]].view(np.uint8).tobytes()[:-1]
        rid = ids[r] if ids else b"rs%d" % r
        out.append(sep.join([b"0", rid, b"0", b"%d" % r]) + sep + body + eol)
    return b"".join(out)


def collapsed(line: bytes) -> bytes:
    """what TpedGenotypeFile hands to addGenotypeRow: 'XY<delim>' triples, digits mapped to letters"""
    f = line.split()
    al = [a.translate(bytes.maketrans(b"1234", b"ACGT")) for a in f[4:]]
    return b"\t".join(a + b for a, b in zip(al[0::2], al[1::2]))


@pytest.mark.parametrize("name", ["cohort_missing", "cohort_complete"])
def test_load_tped_equals_reference_rows(orc, tmp_path, name):
    g = load_golden(name)
    codes = g["codes"]
    M, N = codes.shape
    text = tped_bytes(codes)
    p = tmp_path / "c.tped"
    p.write_bytes(text)
    assert gw.tped_dims(str(p)) == (M, N)
    with gw.GenoStore(M, N) as st:
        assert st.load_tped(str(p)) == M
        assert np.array_equal(st.get_rows(), g["raw_rows"])            # rows of the unmodified reference
    pz = tmp_path / "c.tped.gz"
    with gzip.open(pz, "wb") as f:
        f.write(text)
    assert gw.tped_dims(str(pz)) == (M, N)                             # a .gz needs no rewind
    with gw.GenoStore(M, N) as st:
        assert st.load_tped(str(pz)) == M
        assert np.array_equal(st.get_rows(), g["raw_rows"])
    for path in (p, pz):
        with gw.GenoStore.from_tped(str(path)) as st:
            assert (st.n_snps, st.n_samples) == (M, N)
            assert np.array_equal(st.get_rows(), g["raw_rows"])
            st.select_case_control(g["pheno"])                         # the trimmed table behaves like one created with M rows
            assert np.array_equal(st.marginal_scan()["counts"], g["cc_selected"])
            hits, stats = st.pairwise_scan(30.0)
            assert stats.pairs_tested == M * (M - 1) // 2 and len(hits) == int(g["boost_located"])


@pytest.mark.parametrize("name", ["simple", "cc"])
def test_perl_fixture_files(orc, name):
    path = os.path.join(GOLDEN, f"perl_{name}.tped")
    M, N = gw.tped_dims(path)
    lines = open(path, "rb").read().splitlines()
    assert M == len(lines)
    want = np.stack([orc.pack_text(collapsed(l), N) for l in lines])
    with gw.GenoStore(M, N) as st:
        assert st.load_tped(path) == M
        assert np.array_equal(st.get_rows(), want)


def test_messy_text_and_chunk_boundaries(orc, tmp_path, monkeypatch):
    rng = np.random.default_rng(7)
    M, N = 257, 1003
    codes = rng.choice(4, size=(M, N), p=[0.5, 0.3, 0.15, 0.05]).astype(np.uint8)
    codes[5] = 3                                                        # a row without any call
    codes[6] = 1                                                        # heterozygotes only
    spells = [[b"1 1", b"1 3", b"3 3", b"0 0"], [b"T T", b"T G", b"G G", b"N N"], [b"C C", b"C A", b"A A", b"0 0"]]
    lines = []
    for r in range(M):
        sep = b" " if r % 3 == 0 else b"\t"
        rid = b"rs" + b"x" * int(rng.integers(0, 90)) + b"%d" % r       # marker fields of very different lengths
        line = tped_bytes(codes[r:r + 1], sep=sep, ids=[rid], eol=b"", spell=spells[r % 3])
        if r % 5 == 0:
            line = b"  " + line                                         # trimmed by the reader
        lines.append(line + (b"\r\n" if r % 4 == 0 else b"\n"))
        if r % 50 == 49:
            lines.append(b"\n" if r % 100 == 49 else b" \t\r\n")        # blank lines are skipped
    text = b"".join(lines)[:-1]                                         # last line without a newline
    want = np.stack([orc.pack_text(collapsed(l), N) for l in text.splitlines() if l.strip()])
    p = tmp_path / "m.tped"
    p.write_bytes(text)
    assert gw.tped_dims(str(p)) == (M, N)
    for chunk in (None, 9000, 4097 * 3, 1000):                          # lines straddle chunks, newline blocks, both; 1000: shorter
                                                                        # than a line, the loader has to grow its buffers
        with gw.GenoStore(M, N) as st:
            if chunk:
                st.set_option(gw.OPT_INGEST_CHUNK, chunk)
            assert st.load_tped(str(p)) == M
            assert np.array_equal(st.get_rows(), want)
    with gw.GenoStore.from_tped(str(p)) as st:                          # table sized by the loader itself, one pass
        assert (st.n_snps, st.n_samples) == (M, N)
        assert np.array_equal(st.get_rows(), want)
    # the same through the in-memory entry point, in three calls; an unterminated tail is left to the caller
    with gw.GenoStore(M, N) as st:
        row, pos = 0, 0
        for cut in (len(text) // 3, 2 * len(text) // 3, len(text)):
            n, used = st.put_tped_text(text[pos:cut], first_row=row)
            row, pos = row + n, pos + used
        assert pos < len(text) and b"\n" not in text[pos:]
        n, used = st.put_tped_text(text[pos:] + b"\n", first_row=row)
        assert row + n == M
        assert np.array_equal(st.get_rows(), want)


def test_large_rows_against_oracle_packer(orc, tmp_path):
    rng = np.random.default_rng(11)
    M, N = 1500, 10_000
    codes = rng.choice(4, size=(M, N), p=[0.62, 0.3, 0.07, 0.01]).astype(np.uint8)
    p = tmp_path / "big.tped"
    p.write_bytes(tped_bytes(codes))                                    # 60 MB: several chunks
    with gw.GenoStore(M, N) as st:
        assert st.load_tped(str(p)) == M
        got = st.get_rows()
    assert np.array_equal(got, orc.pack_codes(codes))                   # the pinned packer on AA / AC / CC / 00 in sample order


def test_loader_errors(tmp_path):
    N = 40
    with gw.GenoStore(4, N) as st:
        ok = b"0 rs0 0 0 " + b" ".join([b"A A"] * N) + b"\n"
        third = b"0 rs1 0 1 " + b" ".join([b"A A", b"C C", b"G G"] + [b"A A"] * (N - 3)) + b"\n"
        with pytest.raises(gw.GwasDevError, match="row 1, column 2 introduces a third genotype spelling"):
            st.put_tped_text(ok + third)
        with pytest.raises(gw.GwasDevError, match="fewer than four marker fields"):
            st.put_tped_text(b"0 rs0 0\n")
        with pytest.raises(gw.GwasDevError, match="more genotype lines than the table has rows"):
            st.put_tped_text(ok * 5)
        assert st.put_tped_text(b"") == (0, 0)
        assert st.put_tped_text(b"no newline yet") == (0, 0)
        st.put_tped_text(ok * 4)                                        # the store is still usable
        assert st.call_at(3, N - 1) == "AA"
    with pytest.raises(gw.GwasDevError, match="cannot open"):
        gw.tped_dims(str(tmp_path / "missing.tped"))


def bed_encode(codes):
    """PLINK SNP-major .bed rows: 0 = hom A1, 1 = missing, 2 = het, 3 = hom A2; 4 genotypes per byte, low bits first."""
    M, N = codes.shape
    g = np.array([0, 2, 3, 1], np.uint8)[codes]
    pad = np.zeros((M, (N + 3) // 4 * 4), np.uint8)
    pad[:, :N] = g
    q = pad.reshape(M, -1, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("N", [64, 1001, 4003])
def test_bed_rows_equal_text_rows(orc, tmp_path, N):
    rng = np.random.default_rng(N)
    M = 300
    codes = rng.choice(4, size=(M, N), p=[0.5, 0.3, 0.15, 0.05]).astype(np.uint8)
    codes[3] = 3
    alleles = np.stack([rng.permutation(4)[:2] for _ in range(M)]).astype(np.uint8)
    bed = bed_encode(codes)
    L = "ACGT"
    want = []
    for r in range(M):
        a1, a2 = L[alleles[r, 0]], L[alleles[r, 1]]
        spell = [(a1 + a1).encode(), (a1 + a2).encode(), (a2 + a2).encode(), b"00"]
        want.append(orc.pack_text(b"\t".join(spell[c] for c in codes[r]), N))
    want = np.stack(want)
    with gw.GenoStore(M, N) as st:
        st.put_bed(bed, alleles)
        assert np.array_equal(st.get_rows(), want)
    # file form with the magic number; default alleles A, C == the TPED of the same calls
    p = tmp_path / "c.bed"
    p.write_bytes(bytes([0x6C, 0x1B, 0x01]) + bed.tobytes())
    assert gw.bed_dims(str(p), N) == M
    with gw.GenoStore(M, N) as st, gw.GenoStore(M, N) as st2:
        assert st.load_bed(str(p)) == M
        st2.put_tped_text(tped_bytes(codes))
        assert np.array_equal(st.get_rows(), st2.get_rows())
    bad = tmp_path / "bad.bed"
    bad.write_bytes(bytes([0x6C, 0x1B, 0x00]) + bed.tobytes())
    with pytest.raises(gw.GwasDevError, match="individual-major"):
        gw.bed_dims(str(bad), N)


def test_configs0_from_tped_text_to_statistics(orc, tmp_path):
    """BASELINE configs[0] end to end: 1 000 cases / 1 000 controls x 10 000 SNPs as an 80 MB TPED file -> device loader ->
    case/control selection -> marginal scan, every SNP's counts against the oracle and the allelic chi-square of a
    sample of SNPs (the configuration the reference itself runs on a CPU)."""
    M, N, NCASE = 10_000, 2_000, 1_000
    codes, pheno = orc.simulate(20121127, M, N, NCASE)
    p = tmp_path / "c0.tped"
    p.write_bytes(tped_bytes(codes))
    assert gw.tped_dims(str(p)) == (M, N)
    with gw.GenoStore(M, N) as st:
        assert st.load_tped(str(p)) == M
        rows = st.get_rows()
        assert np.array_equal(rows, orc.pack_codes(codes))
        st.select_case_control(pheno)
        out = st.marginal_scan(mi=False)
    sel, nca, nco = orc.select(rows, N, pheno)
    assert (nca, nco) == (NCASE, N - NCASE)
    assert np.array_equal(out["counts"], orc.cc_counts_selected(sel, nca, nco))
    for r in range(0, M, 97):
        x, pv = orc.chi2_allelic(out["counts"][r, :4], out["counts"][r, 4:])
        s = out["stats"][r]
        assert abs(s["chi2_allelic"] - x) <= 1e-12 * abs(x) and abs(s["p_allelic"] - pv) <= 1e-10 * abs(pv)
