"""The C++ host mirror (DeviceGenoTable + the reference's test-class functions + harness executable) on a
GPU: its printed output must equal what the unmodified reference prints for the same files, and its
per-call virtuals must agree with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "libgwaspp_b200", "gwas_b200")


def run_cli(tped, tfam, flag, tmp_path):
    out = tmp_path / "out.txt"
    r = subprocess.run([CLI, "--tplink", "-g", str(tped), "-p", str(tfam), "--comp-level", "5", flag, "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.rstrip().endswith("DONE")
    return out.read_text()


def write_tplink(tmp_path, codes, pheno):
    txt = {0: "A\tA", 1: "A\tC", 2: "C\tC", 3: "0\t0"}
    tped, tfam = tmp_path / "c.tped", tmp_path / "c.tfam"
    with open(tped, "w") as f:
        for r in range(codes.shape[0]):
            f.write(f"0\trs{r}\t0\t{r}\t" + "\t".join(txt[int(c)] for c in codes[r]) + "\n")
    with open(tfam, "w") as f:
        for i, p in enumerate(pheno):
            f.write(f"Fam_{i}\tInd_{i}\tPat_{i}\tMat_{i}\tx\t{int(p)}\n")
    return tped, tfam


def test_inline_maf_print_equals_reference_output(tmp_path):
    for name in ("simple", "cc"):
        got = run_cli(os.path.join(GOLDEN, f"perl_{name}.tped"), os.path.join(GOLDEN, f"perl_{name}.tfam"),
                      "--test-inline-maf", tmp_path)
        assert got == open(os.path.join(GOLDEN, f"perl_{name}.ref_inline_maf_print.txt")).read()


def test_select_and_inline_cc_maf_shape(tmp_path):
    tped, tfam = os.path.join(GOLDEN, "perl_cc.tped"), os.path.join(GOLDEN, "perl_cc.tfam")
    lines = run_cli(tped, tfam, "--select-cc-maf", tmp_path).splitlines()
    assert lines[0].startswith("-1\t") and lines[0].endswith("s")
    assert [l.split("\t")[0] for l in lines[1:]] == [str(i) for i in range(10)]
    lines = run_cli(tped, tfam, "--inline-cc-maf", tmp_path).splitlines()
    assert [l.split("\t")[0] for l in lines] == [str(i) for i in range(10)]


@pytest.mark.parametrize("name", ["cohort_missing", "cohort_complete"])
def test_compute_boost_output_equals_reference_output(tmp_path, name):
    g = load_golden(name)
    tped, tfam = write_tplink(tmp_path, g["codes"], g["pheno"])
    got = run_cli(tped, tfam, "--test-boost-epi", tmp_path).splitlines()
    ref = str(g["boost_text"]).splitlines()
    # identical except for the elapsed-time line
    assert got[0] == ref[0] == f"Pre-screening {g['codes'].shape[0]} SNP interactions"
    assert got[2:] == ref[2:]
    assert any(l.startswith("Located") for l in got) and len(got) == len(ref)


def test_per_call_virtuals_against_golden(tmp_path):
    g = load_golden("cohort_missing")
    codes, pheno = g["codes"], g["pheno"]
    M, N = codes.shape
    tped, tfam = write_tplink(tmp_path, codes, pheno)
    txt = {0: "AA", 1: "AC", 2: "CC", 3: "00"}
    pair_index = {(int(a), int(b)): k for k, (a, b) in enumerate(g["pairs"])}
    seen = set()
    for line in run_cli(tped, tfam, "--dump-api", tmp_path).splitlines():
        p = line.split()
        tag = p[0]
        assert tag != "MARGIN_MISMATCH"
        seen.add(tag)
        if tag == "call":
            assert p[3] == txt[int(codes[int(p[1]), int(p[2])])]
        elif tag == "whole":
            r = int(p[1])
            assert list(map(int, p[2:5])) == g["whole"][r, :3].tolist() and int(p[5]) == N - int(g["whole"][r, :3].sum())
        elif tag in ("mask_ca", "mask_co", "sel_ca", "sel_co", "mar_ca", "mar_co"):
            r = int(p[1])
            ref = g["cc_masked"] if tag.startswith("mask") else g["cc_selected"]
            off = 0 if tag.endswith("ca") else 4
            assert list(map(int, p[2:6])) == ref[r, off:off + 4].tolist()
        elif tag[0] == "t":
            i, j = int(p[1]), int(p[2])
            if (i, j) in pair_index:
                mode = int(tag[1])
                ref = g[f"tables_mode{mode}"][pair_index[(i, j)]]
                half = ref[:16] if (mode == 0 or tag.endswith("ca")) else ref[16:]
                assert list(map(int, p[3:19])) == half.tolist()
    assert {"call", "whole", "mask_ca", "sel_co", "mar_ca", "t0", "t1_ca", "t2_co", "t3_ca"} <= seen


@pytest.mark.parametrize("name", ["cohort_missing", "cohort_complete"])
def test_contin_and_epi_debug_output_equals_reference_output(orc, tmp_path, name):
    """--contin-debug and --epi-debug (ContingencyDebug / EpistasisDebug, epistasis_func.cpp:84-103, 263-305): the text
    written to the output stream must be the unmodified reference's (tests/golden/make_golden_debug.py); the
    likelihood-ratio values behind --epi-debug / --epi-perform are checked against the pinned pairwise.c restatement."""
    import libgwaspp_b200 as gw
    g, dbg = load_golden(name), load_golden("debug_prints")
    n = int(dbg[name + "_n_snps"])
    codes, pheno = g["codes"][:n], g["pheno"]
    tped, tfam = write_tplink(tmp_path, codes, pheno)
    assert run_cli(tped, tfam, "--contin-debug", tmp_path) == str(dbg[name + "_contin_debug"])
    assert run_cli(tped, tfam, "--epi-debug", tmp_path) == str(dbg[name + "_epi_debug"])
    lines = run_cli(tped, tfam, "--contin-perform", tmp_path).splitlines()
    assert [l.split("\t")[0] for l in lines] == [str(k) for k in range(n * (n - 1) // 2)]
    assert len(run_cli(tped, tfam, "--contin-cc-perform", tmp_path).splitlines()) == n * (n - 1) // 2
    assert run_cli(tped, tfam, "--epi-perform", tmp_path).startswith(f"Case/Control Contingencies {n}: ")
    # gwasdev_epi_pairs == pairwise.c on the 3x3 cores of the mode-1 tables, with EpistasisDebug's fixed set
    N = codes.shape[1]
    ph = np.full(N, 2, np.uint8)
    ph[0:400:2], ph[1:400:2] = 1, 0
    pi, pj = np.triu_indices(n, 1)
    with gw.GenoStore(n, N) as st:
        st.put_rows(orc.pack_codes(codes))
        st.select_case_control(ph)
        tabs = st.pair_tables(pi, pj, 1)
        ll, p = st.epi_pairs(pi, pj, 1)
    core = [0, 1, 2, 4, 5, 6, 8, 9, 10]
    for q in range(len(pi)):
        want = orc.pairwise_epi_test(tabs[q, :16][core], tabs[q, 16:][core])
        assert (np.isnan(want) and np.isnan(ll[q])) or abs(ll[q] - want) <= 1e-12 * abs(want)
        wp = orc.chisq_upper(want, 4)
        assert (np.isnan(wp) and np.isnan(p[q])) or abs(p[q] - wp) <= 1e-10 * abs(wp)


def test_cli_reads_gz_and_bed_genotype_files(tmp_path):
    """The harness hands the genotype file to the device loaders: a gzip-compressed TPED (which the reference's two-pass
    reader cannot rewind) and a SNP-major PLINK .bed give the output of the plain TPED, i.e. the reference's own."""
    import gzip
    from test_gpu_ingest import bed_encode
    g = load_golden("cohort_missing")
    tped, tfam = write_tplink(tmp_path, g["codes"], g["pheno"])
    want = str(g["inline_maf_print"])
    assert run_cli(tped, tfam, "--test-inline-maf", tmp_path) == want
    gz = tmp_path / "c.tped.gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(tped, "rb").read())
    assert run_cli(gz, tfam, "--test-inline-maf", tmp_path) == want
    bed = tmp_path / "c.bed"
    bed.write_bytes(bytes([0x6C, 0x1B, 0x01]) + bed_encode(g["codes"]).tobytes())
    assert run_cli(bed, tfam, "--test-inline-maf", tmp_path) == want
    got = run_cli(bed, tfam, "--test-boost-epi", tmp_path).splitlines()
    assert got[2:] == str(g["boost_text"]).splitlines()[2:]
    # with a .bim next to it the allele letters decide what getCallAt spells (--dump-api prints calls); counts do not move
    with open(tmp_path / "c.bim", "w") as f:
        for r in range(g["codes"].shape[0]):
            f.write(f"1\trs{r}\t0\t{r}\tG\tT\n")
    assert run_cli(bed, tfam, "--test-inline-maf", tmp_path) == want
    spell = {0: "GG", 1: "GT", 2: "TT", 3: "00"}
    calls = [l.split() for l in run_cli(bed, tfam, "--dump-api", tmp_path).splitlines() if l.startswith("call ")]
    assert calls and all(c[3] == spell[int(g["codes"][int(c[1]), int(c[2])])] for c in calls)
    # the reference-shaped host loop (one addGenotypeRow per line) stays available and agrees
    out = tmp_path / "hp.txt"
    r = subprocess.run([CLI, "--tplink", "-g", str(tped), "-p", str(tfam), "--host-parse", "--test-inline-maf", "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and out.read_text() == want
