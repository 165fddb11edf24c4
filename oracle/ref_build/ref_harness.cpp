// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference sources under /root/reference (compiled where they
// lie by oracle/ref_build/Makefile; outputs only into oracle/_ref/).  It lets tests/ and bench.py's
// cpu_baseline / --impl reference legs drive the reference's own classes and test functions:
//   CompressedGenotypeTable{3,4,5}      genetics/genotype/compressed_genotype_table{3,4,5}.{h,cpp}
//   CaseControlSet                       genetics/analyzable/case_control_set.{h,cpp}
//   computeMargins/computeBoost/computeGTest   algorithms/epistasis_func.cpp:706-721,349-506,508-704
//   select_cc_maf/inline_cc_maf/inline_maf_print   algorithms/maf_func.cpp:238-335
//   compute()                            algorithms/computation_engine.cpp:73-86
//   pairwise_epi_test (C)                src/test/pairwise.c:50-133
//   TPED/TFAM readers                    genetics/individual/*_file.cpp (as driven by src/test/gwas_basic.cpp:138-191)
// Nothing here is linked into, imported by, or executed from the product library.
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <fcntl.h>
#include <sys/time.h>
#include <unistd.h>
#include <time.h>

// The harness needs the private table/case-control members of GeneticData and the protected layout
// fields of the tables (to dump packed rows for layout parity).  Access specifiers do not change
// object layout or name mangling, so the reference objects compiled without this remain compatible.
#define private public
#define protected public
#include "libgwaspp.h"
#include "genetics/genetic_data.h"
#include "genetics/genetic_data_file.h"
#include "genetics/genotype/geno_table.h"
#include "genetics/genotype/genotype_tables.h"
#include "genetics/analyzable/case_control_set.h"
#include "genetics/individual/tped_genotype_file.h"
#include "genetics/individual/tfam_phenotype_file.h"
#include "genetics/individual/tfam_annotation_file.h"
#include "algorithms/computation_engine.h"
#include "algorithms/epistasis_func.h"
#include "algorithms/maf_func.h"
#undef private
#undef protected

using namespace libgwaspp::genetics;
using namespace libgwaspp::algorithms;

extern "C" double pairwise_epi_test(int cs[][3], int ct[][3]);  // src/test/pairwise.c:50

namespace {

// Minimal util::indexer (src/util/index_set/indexer.h:39-72): tables only ask for included_size().
class RangeIndexer : public util::indexer {
public:
    explicit RangeIndexer(int n) : n_(n) {}
    int orderOf(const std::string &id) { return atoi(id.c_str()); }
    int indexOf(int ord) { return ord; }
    std::string getIDAtOrderedIndex(int ord) { return std::to_string(ord); }
    void include(int, int) {}
    void include(const std::vector<int> &) {}
    void include(int *, int) {}
    void include(const std::string &, int) {}
    void include(const std::vector<std::string> &) {}
    void exclude(int) {}
    void exclude(const std::vector<int> &) {}
    void exclude(int *, int) {}
    void exclude(const std::string &) {}
    void exclude(const std::vector<std::string> &) {}
    util::IndexIterator *included_begin() { return NULL; }
    util::IndexIterator *included_end() { return NULL; }
    util::IndexIterator *excluded_begin() { return NULL; }
    util::IndexIterator *excluded_end() { return NULL; }
    int included_size() const { return n_; }
    int excluded_size() const { return 0; }
    int maximum_size() const { return n_; }
private:
    int n_;
};

struct Ref {
    GeneticData *gd;
    GenoTable *gt;
    CaseControlSet *ccs;
    marginal_information *margins;
    int n_snps, n_samples, level;
    bool selected;
};

// The reference prints progress to std::cout from constructors and compute(); keep test logs quiet.
struct CoutSilencer {
    std::streambuf *old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

int copy_out(const std::string &s, char *buf, long cap) {
    if (buf && cap > 0) {
        long n = (long)s.size() < cap - 1 ? (long)s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int)s.size();
}

}  // namespace

extern "C" {

int gwasref_sizeof_marginal_information() { return (int)sizeof(marginal_information); }

void *gwasref_create(int n_snps, int n_samples, int level) {
    CoutSilencer q;
    Ref *r = new Ref();
    r->n_snps = n_snps; r->n_samples = n_samples; r->level = level;
    r->gd = new GeneticData((eCompressionLevel)level);
    r->gd->genotyped_markers = new RangeIndexer(n_snps);
    r->gd->genotyped_individs = new RangeIndexer(n_samples);
    r->gd->updateGenotypeTable();               // genetic_data.cpp:46-80 factory switch
    r->gt = r->gd->getGenotypeTable();
    r->gd->ccs = new CaseControlSet(r->gd->genotyped_individs);
    r->ccs = r->gd->ccs;
    r->margins = NULL;
    r->selected = false;
    return r;
}

// Load through the reference's own TPED/TFAM readers exactly as src/test/gwas_basic.cpp:138-191 does.
void *gwasref_load_tplink(const char *tped, const char *tfam, int level) {
    CoutSilencer q;
    Ref *r = new Ref();
    r->level = level;
    r->gd = new GeneticData((eCompressionLevel)level);
    TfamPhenotypeFile ipf;
    TpedGenotypeFile igf;
    TFamAnnotationFile iaf;
    std::string ped(tped), fam(tfam);
    ipf.populateGeneticData(fam, r->gd, '\t');
    igf.populateGeneticData(ped, r->gd, '\t');
    iaf.populateGeneticData(fam, r->gd, '\t');
    r->gt = r->gd->getGenotypeTable();
    r->ccs = r->gd->getCaseControlSet();
    r->n_snps = r->gd->getGenotypedMarkersCount();
    r->n_samples = r->gd->getGenotypedIndividualsCount();
    r->margins = NULL;
    r->selected = false;
    return r;
}

int gwasref_n_snps(void *h) { return ((Ref *)h)->n_snps; }
int gwasref_n_samples(void *h) { return ((Ref *)h)->n_samples; }
int gwasref_n_cases(void *h) { return (int)((Ref *)h)->ccs->getCaseCount(); }
int gwasref_n_controls(void *h) { return (int)((Ref *)h)->ccs->getControlCount(); }

void gwasref_add_row_text(void *h, int r, const char *txt, long len) {
    ((Ref *)h)->gt->addGenotypeRow(r, txt, txt + len, '\t');
}

// codes: one byte per sample, 0 = "AA", 1 = "AC", 2 = "CC", 3 = "00" (missing), 4 = "CA" (reverse het).
void gwasref_add_rows_codes(void *h, int r0, int n_rows, const uint8_t *codes) {
    Ref *r = (Ref *)h;
    static const char *txt[5] = {"AA", "AC", "CC", "00", "CA"};
    std::string line((size_t)r->n_samples * 3, '\t');
    for (int i = 0; i < n_rows; ++i) {
        const uint8_t *c = codes + (size_t)i * r->n_samples;
        for (int s = 0; s < r->n_samples; ++s) {
            line[3 * s] = txt[c[s]][0];
            line[3 * s + 1] = txt[c[s]][1];
        }
        r->gt->addGenotypeRow(r0 + i, line.data(), line.data() + line.size() - 1, '\t');
    }
}

// pheno: one byte per sample, 1 = case, 0 = control, anything else = in neither set.
void gwasref_set_case_control(void *h, const uint8_t *pheno) {
    Ref *r = (Ref *)h;
    std::set<int> ca, co;
    for (int s = 0; s < r->n_samples; ++s) {
        if (pheno[s] == 1) ca.insert(s);
        else if (pheno[s] == 0) co.insert(s);
    }
    r->ccs->reset();
    r->ccs->setCases(ca);
    r->ccs->setControls(co);
}

void gwasref_select(void *h) {
    Ref *r = (Ref *)h;
    r->gt->selectCaseControl(*r->ccs);
    r->selected = true;
}

// whole-cohort counts {aa, ab, bb, xx}  (T5: compressed_genotype_table5.cpp:577-607)
void gwasref_dist(void *h, int row, uint32_t out[4]) {
    GenotypeDistribution d;
    ((Ref *)h)->gt->getGenotypeDistribution(row, d);
    memcpy(out, d.getDistribution()->freq, 16);
}

// mode 0: mask-on-the-fly (T5 :609-657); 1: pre-selected (:659-701); 2: pre-selected + margins (:703-747)
// out = cases {aa,ab,bb,xx} then controls {aa,ab,bb,xx}; mi (may be NULL) receives marginal_information.
void gwasref_cc_dist(void *h, int row, int mode, uint32_t out[8], void *mi) {
    Ref *r = (Ref *)h;
    CaseControlGenotypeDistribution d;
    marginal_information m;
    memset(&m, 0, sizeof m);
    if (mode == 0) r->gt->getCaseControlGenotypeDistribution(row, *r->ccs, d);
    else if (mode == 1) r->gt->getCaseControlGenotypeDistribution(row, d);
    else r->gt->getCaseControlGenotypeDistribution(row, d, m);
    memcpy(out, d.getCaseDistribution()->freq, 16);
    memcpy(out + 4, d.getControlDistribution()->freq, 16);
    if (mi) memcpy(mi, &m, sizeof m);
}

// computeMargins (algorithms/epistasis_func.cpp:706-721) -> copies n_snps * sizeof(marginal_information)
void gwasref_margins(void *h, void *out) {
    Ref *r = (Ref *)h;
    int n = 0;
    int nInd = r->ccs->getCaseCount() + r->ccs->getControlCount();
    computeMargins(*r->gt, nInd, r->margins, n);
    if (out) memcpy(out, r->margins, (size_t)n * sizeof(marginal_information));
}

// mode 0: un-stratified getContingencyTable (T5 :749-800), result in ca only
// mode 1: mask-on-the-fly (:806-895); 2: pre-selected with xx cells (:896-987); 3: margins overload (:989-1150)
void gwasref_pair_table(void *h, int i, int j, int mode, uint32_t ca[16], uint32_t co[16]) {
    Ref *r = (Ref *)h;
    if (mode == 0) {
        ContingencyTable ct;
        r->gt->getContingencyTable(i, j, ct);
        memcpy(ca, ct.getContingencyTable()->contin, 64);
        memset(co, 0, 64);
        return;
    }
    CaseControlContingencyTable t;
    if (mode == 1) r->gt->getCaseControlContingencyTable(i, j, *r->ccs, t);
    else if (mode == 2) r->gt->getCaseControlContingencyTable(i, j, t);
    else r->gt->getCaseControlContingencyTable(i, j, r->margins[i], r->margins[j], t);
    memcpy(ca, t.getCaseContingencyTable()->contin, 64);
    memcpy(co, t.getControlContingencyTable()->contin, 64);
}

// compute(fn, gd, out) for the reference's own test-class entry points (gwas_basic.cpp:195-237).
// which: 0 computeBoost, 1 select_cc_maf, 2 inline_cc_maf, 3 inline_maf_print, 4 genotype_dist_performance
int gwasref_run(void *h, int which, char *buf, long cap) {
    Ref *r = (Ref *)h;
    std::ostringstream out;
    CoutSilencer q;
    switch (which) {
    case 0: compute(computeBoost, r->gd, &out); r->selected = true; break;
    case 1: compute(select_cc_maf, r->gd, &out); r->selected = true; break;
    case 2: compute(inline_cc_maf, r->gd, &out); break;
    case 3: compute(inline_maf_print, r->gd, &out); break;
    case 4: compute(genotype_dist_performance, r->gd, &out); break;
    case 5: case 6: {   // ContingencyDebug / EpistasisDebug (epistasis_func.cpp:84-103, 263-305): only inp.gd is read
        std::set<std::string> mids, iids;
        BasicInput bi(r->gd, &mids, &iids);
        IndexedInput ii(&bi);
        if (which == 5) ContingencyDebug(&ii, &out);
        else {          // its log-likelihood lines go to stdout through printf: keep them out of the test log
            fflush(stdout);
            const int keep = dup(1), nul = open("/dev/null", O_WRONLY);
            dup2(nul, 1);
            EpistasisDebug(&ii, &out);
            fflush(stdout);
            dup2(keep, 1);
            close(keep); close(nul);
        }
        break;
    }
    default: return -1;
    }
    return copy_out(out.str(), buf, cap);
}

// computeGTest (epistasis_func.cpp:508-704) on caller-given pairs. Needs gwasref_select + gwasref_margins.
void gwasref_gtest(void *h, int n, const uint32_t *pi, const uint32_t *pj, double *stat, double *z) {
    Ref *r = (Ref *)h;
    std::vector<SNPInteractionPair> v;
    std::vector<double> zv;
    for (int k = 0; k < n; ++k) v.push_back(SNPInteractionPair(SNPPair(pi[k], pj[k]), 0.0));
    computeGTest(*r->gt, r->margins, r->ccs->getCaseCount() + r->ccs->getControlCount(), v, zv);
    for (int k = 0; k < n; ++k) { stat[k] = v[k].second; z[k] = zv[k]; }
}

// src/test/pairwise.c:50-133 on dense 3x3 tables; also returns pchisq(ll, 4, 0, 0) as in :44.
double gwasref_pairwise_c(const int cs[9], const int ct[9], double *pval) {
    int a[3][3], b[3][3];
    memcpy(a, cs, sizeof a);
    memcpy(b, ct, sizeof b);
    double ll = pairwise_epi_test(a, b);
    if (pval) *pval = pchisq(ll, 4.0, 0, 0);
    return ll;
}

// Packed-layout probes (for layout parity of the restatement).
// raw row: blocks_per_row ushorts (T5: [hdr][plane1][plane2], compressed_genotype_table5.cpp:34-153)
int gwasref_raw_row(void *h, int row, uint16_t *out, int cap) {
    Ref *r = (Ref *)h;
    int n = (int)r->gt->blocks_per_row;
    if (r->gt->data == NULL) return 0;          // a table without host storage (the device table, level 6)
    if (out) memcpy(out, r->gt->data + (size_t)row * n, 2 * (size_t)(n < cap ? n : cap));
    return n;
}
// compacted row: nCaseControlBlockCount ushorts (T5: compressed_genotype_table5.cpp:443-575)
int gwasref_selected_row(void *h, int row, uint16_t *out, int cap, int geom[4]) {
    Ref *r = (Ref *)h;
    if (r->gt->m_cases_controls == NULL) return 0;
    int n = (int)r->gt->nCaseControlBlockCount;
    if (geom) {
        geom[0] = r->gt->nCaseBlockCount; geom[1] = r->gt->nControlBlockCount;
        geom[2] = r->gt->nControlBlockOffset; geom[3] = n;
    }
    if (out) memcpy(out, r->gt->m_cases_controls + (size_t)row * n, 2 * (size_t)(n < cap ? n : cap));
    return n;
}
// decoded call at (row, col) via the table's own operator() + decodeGenotype (geno_table.h:59-60)
void gwasref_call_at(void *h, int row, int col, char out[3]) {
    const char *s = ((Ref *)h)->gt->getCallAt(row, col);
    out[0] = s[0]; out[1] = s[1]; out[2] = 0;
}

// Wall-clock of the reference's own phases on this host, for the CPU baseline.
// phase 0: selectCaseControl; 1: computeMargins; 2: per-SNP pre-selected CC counts + MinorAlleleFrequency
// (the select_cc_maf loop body, maf_func.cpp:256-267, without its per-item timer/stream writes);
// 3: compute(computeBoost) whole call (pre-screen + G-test + print).
double gwasref_time_phase(void *h, int phase, int reps) {
    Ref *r = (Ref *)h;
    CoutSilencer q;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int rep = 0; rep < reps; ++rep) {
        if (phase == 0) { r->gt->selectCaseControl(*r->ccs); r->selected = true; }
        else if (phase == 1) {
            int n = 0;
            computeMargins(*r->gt, r->ccs->getCaseCount() + r->ccs->getControlCount(), r->margins, n);
        } else if (phase == 2) {
            CaseControlGenotypeDistribution d;
            double tot, maf, acc = 0;
            for (int i = 0; i < r->n_snps; ++i) {
                r->gt->getCaseControlGenotypeDistribution(i, d);
                MinorAlleleFrequency(*d.getCaseDistribution(), tot, maf); acc += maf;
                MinorAlleleFrequency(*d.getControlDistribution(), tot, maf); acc += maf;
            }
            if (acc == -1.0) printf("x");
        } else {
            std::ostringstream out;
            compute(computeBoost, r->gd, &out);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

}  // extern "C"
