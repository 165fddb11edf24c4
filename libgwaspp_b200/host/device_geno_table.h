// DeviceGenoTable: the reference's GenoTable interface (genetics/genotype/geno_table.h:47-87 with
// SingleMarkerAnalyzable single_marker_analyzable.h:122-145, PairwiseMarkerAnalyzable
// pairwise_marker_analyzable.h:165-187 and CaseControlSelectable case_control_selectable.h:38-56) for the
// association path, backed by the device-resident store behind the C-ABI in include/gwasdev.h.
//
// Same method names, argument meaning and error behaviour (assert/abort, no exceptions, no return codes)
// as the reference's CompressedGenotypeTable5. Per-call virtuals fetch one row or one pair from the
// device (slow, like any per-item PCIe round trip); the batch entry points at the bottom are what the
// test-class functions in test_functions.cpp use.
#pragma once
#include <string>
#include <vector>

#include "gwas_types.h"
#include "gwasdev.h"

namespace libgwaspp {
namespace genetics {

class DeviceGenoTable {
public:
    DeviceGenoTable(int n_markers, int n_individuals, int device = 0);
    explicit DeviceGenoTable(const std::string &tped_path, int device = 0);   // dimensions and rows from the file, one pass
    virtual ~DeviceGenoTable();

    // ---- util::Table<DataBlock> (src/util/table/table.h:48-74)
    int row_size() const { return max_row; }
    int column_size() const { return max_column; }
    DataBlock operator()(int r, int c);

    // ---- GenoTable loading / codec
    void addGenotype(int rIdx, int cIdx, const std::string &gt);
    void addGenotypeRow(int rIdx, std::string::const_iterator &it, std::string::const_iterator &it_end, char delim);
    void addGenotypeRow(int rIdx, const char *p_begin, const char *p_end, char delim);
    ushort encodeGenotype(const std::string &gt);
    const char *decodeGenotype(ushort encoded_gt);
    bool isGenotypeHomozygous(ushort enc);
    const char *getCallAt(uint marker_idx, uint individ_idx) { return decodeGenotype(getGenotypeAt(marker_idx, individ_idx)); }
    DataBlock getGenotypeAt(uint marker_idx, uint individ_idx) { return (*this)((int)marker_idx, (int)individ_idx); }
    uint getPossibleGenotypeCount() const { return 17; }

    // ---- CaseControlSelectable
    void selectCaseControl(CaseControlSet &ccs);

    // ---- SingleMarkerAnalyzable
    void selectMarker(uint rIdx);
    void getGenotypeDistribution(uint rIdx, GenotypeDistribution &dist);
    void getCaseControlGenotypeDistribution(uint rIdx, CaseControlSet &ccs, CaseControlGenotypeDistribution &ccgd);
    void getCaseControlGenotypeDistribution(uint rIdx, CaseControlGenotypeDistribution &ccgd);
    void getCaseControlGenotypeDistribution(uint rIdx, CaseControlGenotypeDistribution &ccgd, marginal_information &m);

    // ---- PairwiseMarkerAnalyzable
    void selectMarkerPair(uint rIdx1, uint rIdx2);
    void getContingencyTable(uint rIdx1, uint rIdx2, ContingencyTable &ct);
    void getContingencyTable(uint rIdx1, uint rIdx2, ushort *column_set, ContingencyTable &ct);
    void getCaseControlContingencyTable(uint rIdx1, uint rIdx2, CaseControlSet &ccs, CaseControlContingencyTable &ccct);
    void getCaseControlContingencyTable(uint rIdx1, uint rIdx2, CaseControlContingencyTable &ccct);
    void getCaseControlContingencyTable(uint rIdx1, uint rIdx2, const marginal_information &m1,
                                        const marginal_information &m2, CaseControlContingencyTable &ccct);

    // ---- batch entry points (no reference counterpart: one call instead of a host loop)
    // whole genotype file parsed, labelled and packed on the device (rows from first_row on): what the reference's
    // TpedGenotypeFile::parseNextGenotypeRecord + addGenotypeRow loop does line by line on the host
    // (genetics/individual/individual_genotype_file.cpp:63-106, tped_genotype_file.cpp:110-185). Returns rows loaded.
    int loadTransposedPlink(const std::string &tped_path, int first_row = 0);
    // PLINK .bed (SNP-major); alleles = 2 per row, indices into "ACGT" of A1 / A2 (may be empty: A, C)
    int loadBed(const std::string &bed_path, const std::vector<unsigned char> &alleles = std::vector<unsigned char>(), int first_row = 0);
    void flush();                                                       // push buffered rows to the device
    void computeMargins(std::vector<marginal_information> &out);        // == algorithms::computeMargins over all rows
    void scanCaseControl(std::vector<frequency_table> &cases, std::vector<frequency_table> &controls,
                         std::vector<gwasdev_snp_stats> *stats = nullptr);
    // n > 1: screenPairs runs on n devices of this process (this table's own and n - 1 replicas made over NVLink on first
    // use), sharded by tile pairs inside the library, hits combined with an NCCL all-gather (gwasdev_pairwise_scan_multi)
    void useDevices(int n);
    int deviceCount() const { return n_devices; }
    void screenPairs(double threshold, std::vector<gwasdev_hit> &hits, gwasdev_pair_stats *stats = nullptr,
                     uint shard = 0, uint n_shards = 1);
    void gtestPairs(const std::vector<gwasdev_hit> &hits, std::vector<double> &stat, std::vector<double> &z);
    gwasdev_store *handle() { flush(); return store; }

private:
    void ensureMasks(CaseControlSet &ccs);
    int max_row, max_column;
    gwasdev_store *store;
    uint plane_blocks;
    std::vector<ushort> pending;          // packed rows waiting for upload
    int pending_first, pending_count;
    std::vector<ushort> cell_row;         // scratch row for addGenotype
    uint64_t selected_rev;                // CaseControlSet revision currently compacted on the device
    const CaseControlSet *selected_set;
    uint64_t fly_rev;                     // CaseControlSet whose masks the mask-on-the-fly overloads currently use
    const CaseControlSet *fly_set;
    char call_buf[3];
    int device0, n_devices;
    std::vector<gwasdev_store *> replicas;   // the same table on the other devices (multi-device screen)
    uint64_t replica_rev;                  // selection revision the replicas carry (0: stale, rebuild)
    void dropReplicas();
    void syncReplicas();
};

}  // namespace genetics
}  // namespace libgwaspp
