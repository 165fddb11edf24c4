// Host-side C++ mirror of the reference's result carriers and PODs for the association path.
// Same names, members and accessor signatures as the reference so that code written against
// libgwaspp's GenoTable compiles against DeviceGenoTable unchanged; the bodies are this repo's own.
//
//   frequency_table / marginal_information / CONTIN_TABLE_T   genetics/genotype/common_genotype.h:67-75,101-106,182-192
//   GenotypeDistribution / CaseControlGenotypeDistribution      genetics/genotype/single_marker_analyzable.h:41,81
//   ContingencyTable / CaseControlContingencyTable              genetics/genotype/pairwise_marker_analyzable.h:44,103
//   CaseControlSet                                              genetics/analyzable/case_control_set.h:46-95
#pragma once
#include <cstdint>
#include <cstring>
#include <set>
#include <vector>

typedef unsigned char byte;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;
typedef ushort DataBlock;

namespace libgwaspp {
namespace genetics {

constexpr int GENOTYPE_COUNT = 4;

union frequency_table {          // {aa, ab, bb, xx}: first-seen homozygote, heterozygote, second homozygote, missing
    uint freq[GENOTYPE_COUNT];
    struct { uint aa, ab, bb, xx; };
};

union header_table {             // genotype encodings (4*idx(c1)+idx(c2) over "ACGT") behind codes xx/aa/ab/bb; 0xFFFF = unseen
    ulong l;
    ushort header[4];
    struct { ushort xx, aa, ab, bb; };
};

struct marginal_information {    // byte-compatible with gwasdev_marginal_information (192 bytes)
    frequency_table margins, cases, controls;
    double dMarginalEntropy, dMarginalEntropy_Y;
    double dPbc[2 * GENOTYPE_COUNT];   // P(genotype | class): cases[4], controls[4]
    double dPca[2 * GENOTYPE_COUNT];   // P(class | genotype): cases[4], controls[4]
};
static_assert(sizeof(marginal_information) == 192, "marginal_information layout");

union extended_contingency_table {   // 4x4 row-major, rows A in {AA, Aa, aa, xx}, columns B likewise
    uint contin[16];
    struct { uint AA_BB, AA_Bb, AA_bb, AA_xx, Aa_BB, Aa_Bb, Aa_bb, Aa_xx, aa_BB, aa_Bb, aa_bb, aa_xx, xx_BB, xx_Bb, xx_bb, xx_xx; };
};
typedef extended_contingency_table CONTIN_TABLE_T;

class GenotypeDistribution {
public:
    GenotypeDistribution() { reset(); }
    virtual ~GenotypeDistribution() {}
    uint getCurrentIndex() { return current_rIdx; }
    const frequency_table *getDistribution() { return &distribution; }
    const header_table *getGenotypes() { return &genotypes; }
    void setDistribution(const frequency_table &ft) { distribution = ft; }
    void setGenotypes(const header_table &h) { genotypes = h; }
    void setCurrentIndex(uint r) { current_rIdx = r; }
    void reset() { memset(&distribution, 0, sizeof distribution); genotypes.l = 0xFFFF000000000000ul; current_rIdx = (uint)-1; }
protected:
    uint current_rIdx;
    frequency_table distribution;
    header_table genotypes;
};

class CaseControlGenotypeDistribution {
public:
    CaseControlGenotypeDistribution() { reset(); }
    virtual ~CaseControlGenotypeDistribution() {}
    uint getCurrentIndex() { return current_rIdx; }
    const frequency_table *getCaseDistribution() { return &case_dist; }
    const frequency_table *getControlDistribution() { return &control_dist; }
    const header_table *getGenotypes() { return &genotypes; }
    void setCaseDistribution(const frequency_table &ft) { case_dist = ft; }
    void setControlDistribution(const frequency_table &ft) { control_dist = ft; }
    void setCurrentIndex(uint r) { current_rIdx = r; }
    void reset() { memset(&case_dist, 0, sizeof case_dist); memset(&control_dist, 0, sizeof control_dist); genotypes.l = 0xFFFF000000000000ul; current_rIdx = (uint)-1; }
protected:
    uint current_rIdx;
    frequency_table case_dist, control_dist;
    header_table genotypes;
};

class ContingencyTable {
public:
    ContingencyTable() { reset(); }
    virtual ~ContingencyTable() {}
    uint getMarkerAIndex() { return ma_rIdx; }
    uint getMarkerBIndex() { return mb_rIdx; }
    void setMarkerAIndex(uint r) { ma_rIdx = r; }
    void setMarkerBIndex(uint r) { mb_rIdx = r; }
    const CONTIN_TABLE_T *getContingencyTable() { return &cont; }
    const header_table *getMarkerAHeader() { return &ma_header; }
    const header_table *getMarkerBHeader() { return &mb_header; }
    void setContingency(const CONTIN_TABLE_T &ct) { cont = ct; }
    void setHeaderA(const header_table &h) { ma_header = h; }
    void setHeaderB(const header_table &h) { mb_header = h; }
    void reset() { memset(&cont, 0, sizeof cont); ma_header.l = mb_header.l = 0xFFFF000000000000ul; ma_rIdx = mb_rIdx = (uint)-1; }
protected:
    uint ma_rIdx, mb_rIdx;
    CONTIN_TABLE_T cont;
    header_table ma_header, mb_header;
};

class CaseControlContingencyTable {
public:
    CaseControlContingencyTable() { reset(); }
    virtual ~CaseControlContingencyTable() {}
    uint getMarkerAIndex() { return ma_rIdx; }
    uint getMarkerBIndex() { return mb_rIdx; }
    void setMarkerAIndex(uint r) { ma_rIdx = r; }
    void setMarkerBIndex(uint r) { mb_rIdx = r; }
    const header_table *getMarkerAHeader() { return &ma_header; }
    const header_table *getMarkerBHeader() { return &mb_header; }
    const CONTIN_TABLE_T *getCaseContingencyTable() { return &case_contin; }
    const CONTIN_TABLE_T *getControlContingencyTable() { return &control_contin; }
    void setCaseContingency(const CONTIN_TABLE_T &ct) { case_contin = ct; }
    void setControlContingency(const CONTIN_TABLE_T &ct) { control_contin = ct; }
    void updateContingencyTables(const CONTIN_TABLE_T &cs, const CONTIN_TABLE_T &ct) { case_contin = cs; control_contin = ct; }
    void setHeaderA(const header_table &h) { ma_header = h; }
    void setHeaderB(const header_table &h) { mb_header = h; }
    void reset() { memset(&case_contin, 0, sizeof case_contin); memset(&control_contin, 0, sizeof control_contin); ma_header.l = mb_header.l = 0xFFFF000000000000ul; ma_rIdx = mb_rIdx = (uint)-1; }
protected:
    uint ma_rIdx, mb_rIdx;
    CONTIN_TABLE_T case_contin, control_contin;
    header_table ma_header, mb_header;
};

// Case / control membership over genotyped-individual order. Only the 1-bit stream masks are kept: they
// are what the bit-plane tables consume (the reference's 2-bit replicated masks serve its 2-bit block
// table only).
class CaseControlSet {
public:
    explicit CaseControlSet(int n_individuals);
    uint getMaximumIndex() const { return max_index; }
    uint getCaseCount() const { return case_count; }
    uint getControlCount() const { return ctrl_count; }
    uint getTotalCount() const { return case_count + ctrl_count; }
    void setCases(const std::set<int> &case_idx);
    void setControls(const std::set<int> &ctrl_idx);
    void setAllAsCases();
    void setAllAsControls();
    void reset();
    const ushort *stream_case_begin() { return stream_case_set.data(); }
    const ushort *stream_case_end() { return stream_case_set.data() + stream_case_set.size(); }
    const ushort *stream_control_begin() { return stream_control_set.data(); }
    const ushort *stream_control_end() { return stream_control_set.data() + stream_control_set.size(); }
    bool isCase(uint idx) const { return (stream_case_set[idx >> 4] >> (idx & 15)) & 1; }
    bool isControl(uint idx) const { return (stream_control_set[idx >> 4] >> (idx & 15)) & 1; }
    uint64_t revision() const { return rev; }   // bumped on every change; lets the table skip re-uploading masks
private:
    uint max_index, case_count, ctrl_count;
    uint64_t rev;
    std::vector<ushort> stream_case_set, stream_control_set;
};

}  // namespace genetics
}  // namespace libgwaspp
