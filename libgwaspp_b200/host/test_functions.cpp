#include "test_functions.h"

#include <cassert>
#include <cstdio>
#include <iostream>
#include <time.h>

namespace libgwaspp {
namespace genetics {

GeneticData::GeneticData(int m, int n, int device)
    : n_markers(m), n_individs(n), geno_tbl(new DeviceGenoTable(m, n, device)), ccs(new CaseControlSet(n)) {}
GeneticData::~GeneticData() { delete geno_tbl; delete ccs; }

void GeneticData::setCaseControlSet(const std::set<int> &cases, const std::set<int> &controls) {
    ccs->reset();
    std::set<int> co(controls);
    for (int c : cases)                      // an index in both sets stays a case (genetic_data.cpp:149-154)
        if (co.erase(c)) std::cout << "Index: " << c << " found in both Case/Controls; Removing from controls" << std::endl;
    ccs->setCases(cases);
    ccs->setControls(co);
}

}  // namespace genetics

namespace algorithms {

namespace {
struct Lap {
    timespec t0;
    Lap() { clock_gettime(CLOCK_MONOTONIC, &t0); }
    double seconds() const { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec); }
};
// "%lu.%09lus" as PRINT_LAPSE does with NANO_TIME (src/util/time/timing.h:45-52,100-106)
void print_lapse(std::ostream &out, const char *prefix, double s) {
    char buf[64];
    unsigned long sec = (unsigned long)s, ns = (unsigned long)((s - (double)sec) * 1e9);
    snprintf(buf, sizeof buf, "%lu.%09lus", sec, ns);
    out << prefix << buf;
}
}  // namespace

void compute(void (*f)(GeneticData *, std::ostream *), GeneticData *gd, std::ostream *out) {
    Lap lap;
    if (out == NULL) out = &std::cout;
    f(gd, out);
    print_lapse(std::cout, "Total runtime: ", lap.seconds());
    std::cout << std::endl;
}

void compute_maf_perform(GeneticData *gd, std::ostream *) {
    const int M = gd->getGenotypedMarkersCount();
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    std::vector<uint> c((size_t)M * 4);
    int rc = gwasdev_counts(gt.handle(), 0, (uint64_t)M, 0, c.data());
    assert(rc == GWASDEV_OK); (void)rc;
    double tot, maf;
    for (int i = 0; i < M; ++i) { frequency_table ft; memcpy(ft.freq, &c[4 * (size_t)i], 16); MinorAlleleFrequency(ft, tot, maf); }
}

void select_cc_maf(GeneticData *gd, std::ostream *out) {
    const int M = gd->getGenotypedMarkersCount();
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    Lap sel;
    gt.selectCaseControl(*gd->getCaseControlSet());
    *out << (int)-1;
    print_lapse(*out, "\t", sel.seconds());
    *out << std::endl;
    Lap scan;
    std::vector<frequency_table> ca, co;
    gt.scanCaseControl(ca, co);
    const double per = scan.seconds() / (M > 0 ? M : 1);      // one batched scan: report its per-SNP share
    double tot, maf;
    for (int i = 0; i < M; ++i) {
        *out << (int)i;
        print_lapse(*out, "\t", per);
        *out << std::endl;
        MinorAlleleFrequency(ca[i], tot, maf);
        MinorAlleleFrequency(co[i], tot, maf);
    }
}

void inline_cc_maf(GeneticData *gd, std::ostream *out) {
    const int M = gd->getGenotypedMarkersCount();
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    CaseControlSet &ccs = *gd->getCaseControlSet();
    Lap scan;
    CaseControlGenotypeDistribution first;
    if (M > 0) gt.getCaseControlGenotypeDistribution(0, ccs, first);     // uploads the masks when they changed
    std::vector<uint> c((size_t)M * 8);
    int rc = gwasdev_counts(gt.handle(), 0, (uint64_t)M, 1, c.data());
    assert(rc == GWASDEV_OK); (void)rc;
    const double per = scan.seconds() / (M > 0 ? M : 1);
    double tot, maf;
    for (int i = 0; i < M; ++i) {
        *out << (int)i;
        print_lapse(*out, "\t", per);
        *out << std::endl;
        frequency_table a, b;
        memcpy(a.freq, &c[8 * (size_t)i], 16); memcpy(b.freq, &c[8 * (size_t)i + 4], 16);
        MinorAlleleFrequency(a, tot, maf);
        MinorAlleleFrequency(b, tot, maf);
    }
}

void inline_maf_print(GeneticData *gd, std::ostream *out) {
    const int M = gd->getGenotypedMarkersCount(), N = gd->getGenotypedIndividualsCount();
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    std::vector<uint> c((size_t)M * 4);
    int rc = gwasdev_counts(gt.handle(), 0, (uint64_t)M, 0, c.data());
    assert(rc == GWASDEV_OK); (void)rc;
    double tot, maf;
    for (int i = 0; i < M; ++i) {
        frequency_table ft;
        memcpy(ft.freq, &c[4 * (size_t)i], 16);
        MinorAlleleFrequency(ft, tot, maf);
        *out << (int)(N - tot) << "\t" << ft.aa << "\t" << ft.ab << "\t" << ft.bb << std::endl;
    }
}

void genotype_dist_performance(GeneticData *gd, std::ostream *out) {
    const int M = gd->getGenotypedMarkersCount();
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    Lap scan;
    std::vector<uint> c((size_t)M * 4);
    int rc = gwasdev_counts(gt.handle(), 0, (uint64_t)M, 0, c.data());
    assert(rc == GWASDEV_OK); (void)rc;
    const double per = scan.seconds() / (M > 0 ? M : 1);
    for (int i = 0; i < M; ++i) { *out << (int)i; print_lapse(*out, "\t", per); *out << std::endl; }
}

void computeMargins(DeviceGenoTable &gt, int nIndivids, marginal_information *&pMargins, int &nMarkerCount) {
    (void)nIndivids;
    nMarkerCount = gt.row_size();
    if (pMargins != NULL) delete[] pMargins;
    pMargins = new marginal_information[nMarkerCount];
    std::vector<marginal_information> tmp;
    gt.computeMargins(tmp);
    memcpy(pMargins, tmp.data(), sizeof(marginal_information) * (size_t)nMarkerCount);
}

void computeGTest(DeviceGenoTable &gt, marginal_information *, uint, std::vector<SNPInteractionPair> &passingThreshold,
                  std::vector<double> &zval) {
    std::vector<gwasdev_hit> hits(passingThreshold.size());
    for (size_t k = 0; k < hits.size(); ++k) { hits[k].i = passingThreshold[k].first.first; hits[k].j = passingThreshold[k].first.second; hits[k].stat = passingThreshold[k].second; }
    std::vector<double> stat;
    gt.gtestPairs(hits, stat, zval);
    for (size_t k = 0; k < hits.size(); ++k) passingThreshold[k].second = stat[k];
}

void computeBoost(GeneticData *gd, std::ostream *out) {
    CaseControlSet &ccs = *gd->getCaseControlSet();
    const int nIndivids = (int)(ccs.getCaseCount() + ccs.getControlCount());
    DeviceGenoTable &gt = *gd->getGenotypeTable();
    gt.selectCaseControl(ccs);
    const double thresholdRecord = 30.0;
    *out << "Pre-screening " << gt.row_size() << " SNP interactions" << std::endl;
    Lap lap;
    std::vector<gwasdev_hit> hits;
    gt.screenPairs(thresholdRecord, hits);                // margins + exhaustive screen + exact re-score on the device
    print_lapse(*out, "", lap.seconds());
    *out << std::endl;
    std::vector<SNPInteractionPair> passingThreshold;
    for (const gwasdev_hit &h : hits) passingThreshold.push_back(SNPInteractionPair(SNPPair(h.i, h.j), h.stat));
    *out << "Located " << passingThreshold.size() << " potential interactions" << std::endl;
    *out << "Performing deeper analysis of SNPs" << std::endl;
    std::vector<double> zval;
    computeGTest(gt, NULL, (uint)nIndivids, passingThreshold, zval);
    int idx = 0;
    char buf[256];
    for (size_t k = 0; k < passingThreshold.size(); ++k) {
        if (passingThreshold[k].second > thresholdRecord) {
            snprintf(buf, sizeof buf, "%7d\t%7d\t%7d\t%f\t%f\t%f\t%f", idx++, (int)passingThreshold[k].first.first,
                     (int)passingThreshold[k].first.second, 0.0, 0.0, passingThreshold[k].second, zval[k]);
            *out << buf << std::endl;
        }
    }
}

}  // namespace algorithms
}  // namespace libgwaspp
