// Host mirror of the reference's test-class API for the association path:
//   compute(fn, gd, out)                                   algorithms/computation_engine.h:72-73, .cpp:73-86
//   select_cc_maf / inline_cc_maf / inline_maf_print /
//   genotype_dist_performance / compute_maf_perform        algorithms/maf_func.h:58-74, maf_func.cpp:224-335
//   MinorAlleleFrequency                                   algorithms/maf_func.h:46-54
//   computeMargins / computeBoost / computeGTest           algorithms/epistasis_func.h:66-83, .cpp:349-721
// Same names, signatures (GeneticData*, ostream*) and printed formats; each function makes one batched
// call into the device library instead of a per-item host loop.
#pragma once
#include <ostream>
#include <set>
#include <utility>
#include <vector>

#include "device_geno_table.h"

namespace libgwaspp {
namespace genetics {

// Stand-in for the reference's GeneticData facade (genetics/genetic_data.h:79-158): owns the table and
// the case/control set and exposes the accessors the test functions use.
class GeneticData {
public:
    GeneticData(int n_markers, int n_individuals, int device = 0);
    GeneticData(const std::string &tped_path, int device);   // genotype table sized from and loaded with the file
    ~GeneticData();
    int getGenotypedIndividualsCount() const { return n_individs; }
    int getGenotypedMarkersCount() const { return n_markers; }
    void addGenotypeRow(int r, const char *p_begin, const char *p_end, char delim) { geno_tbl->addGenotypeRow(r, p_begin, p_end, delim); }
    const char *getGenotype(int r, int c) { return geno_tbl->getCallAt((uint)r, (uint)c); }
    DeviceGenoTable *getGenotypeTable() { return geno_tbl; }
    CaseControlSet *getCaseControlSet() { return ccs; }
    void setCaseControlSet(const std::set<int> &cases, const std::set<int> &controls);
private:
    int n_markers, n_individs;
    DeviceGenoTable *geno_tbl;
    CaseControlSet *ccs;
};

}  // namespace genetics

namespace algorithms {

using namespace libgwaspp::genetics;

typedef std::pair<uint, uint> SNPPair;
typedef std::pair<SNPPair, double> SNPInteractionPair;

inline void MinorAlleleFrequency(const frequency_table &ft, double &tot, double &maf) {
    tot = ft.aa; maf = 2.0 * tot; tot += ft.ab; maf += ft.ab; tot += ft.bb; maf /= tot;
    if (maf < 0.5) maf = 1.0 - maf;     // the reference returns max(f, 1 - f) under this name
}

void compute(void (*f)(GeneticData *, std::ostream *), GeneticData *gd, std::ostream *out);

void compute_maf_perform(GeneticData *gd, std::ostream *out);
void select_cc_maf(GeneticData *gd, std::ostream *out);
void inline_cc_maf(GeneticData *gd, std::ostream *out);
void inline_maf_print(GeneticData *gd, std::ostream *out);
void genotype_dist_performance(GeneticData *gd, std::ostream *out);

// --contin-debug / --contin-perform / --contin-cc-perform / --epi-debug / --epi-perform
// (algorithms/epistasis_func.cpp:84-103, 105-135, 137-202, 263-305, 307-347). The reference passes an IndexedInput as
// void*; only its GeneticData is read, so these take the GeneticData directly. Pairs are enumerated in the reference's
// order and evaluated in batches on the device.
void ContingencyDebug(GeneticData *gd, std::ostream *out);
void ContingencyPerformance(GeneticData *gd, std::ostream *out);
void ContingencyCCPerformance(GeneticData *gd, std::ostream *out);
void EpistasisDebug(GeneticData *gd, std::ostream *out);
void EpistasisPerformance(GeneticData *gd, std::ostream *out);
void printContingencyTable(const CONTIN_TABLE_T &ct, std::ostream &out);

void computeMargins(DeviceGenoTable &gt, int nIndivids, marginal_information *&pMargins, int &nMarkerCount);
void computeGTest(DeviceGenoTable &gt, marginal_information *pMargins, uint nIndivids,
                  std::vector<SNPInteractionPair> &passingThreshold, std::vector<double> &zval);
void computeBoost(GeneticData *gd, std::ostream *out);

}  // namespace algorithms
}  // namespace libgwaspp
