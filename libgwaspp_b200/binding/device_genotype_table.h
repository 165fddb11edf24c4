// DeviceGenotypeTable -- libgwaspp's GenoTable on a B200.
//
// This is the file a libgwaspp maintainer adds next to genetics/genotype/compressed_genotype_table5.h: a subclass of the
// reference's own abstract GenoTable (genetics/genotype/geno_table.h:47-87 = util::Table<DataBlock> src/util/table/table.h:48-74
// + SingleMarkerAnalyzable single_marker_analyzable.h:122-145 + PairwiseMarkerAnalyzable pairwise_marker_analyzable.h:165-187
// + CaseControlSelectable case_control_selectable.h:38-56) whose storage and arithmetic live behind the C-ABI of
// include/gwasdev.h. It is compiled against the reference's headers (it includes them, it copies nothing from them) and
// links libgwasdev.so; with the two-line factory case in GeneticData::updateGenotypeTable (genetics/genetic_data.cpp:60-79,
// INTEGRATION.md) the reference's unmodified readers, compute() and test functions -- inline_maf_print, select_cc_maf,
// inline_cc_maf, computeMargins, computeBoost, computeGTest, Contingency*/Epistasis* -- run on the device table.
//
// The reference's test functions call the per-item virtuals in loops (one row, or one pair, per call). A device round
// trip per item would cost ~20 us each, so every overload serves its items from a host-side block fetched with ONE
// batched C-ABI call: rows [r, r + n) for the per-row overloads, pairs (i, j .. j + n - 1) for the per-pair overloads --
// the order in which computeMargins / computeBoost / ContingencyDebug walk them -- with n = 64 for an isolated request
// (computeGTest's hits) and doubling up to 8192 rows / 4096 pairs while the caller keeps walking. Results are those of
// the per-item calls.
#ifndef DEVICE_GENOTYPE_TABLE_H
#define DEVICE_GENOTYPE_TABLE_H

#include <vector>

#include "genetics/genotype/geno_table.h"
#include "gwasdev.h"

namespace libgwaspp {
namespace genetics {

class DeviceGenotypeTable : public GenoTable {
public:
    DeviceGenotypeTable( util::indexer *markers, util::indexer *individs, int device = 0 );

    DataBlock operator()( int r, int c );

    void addGenotype( int rIdx, int cIdx, const string &gt );
    void addGenotypeRow( int rIdx, string::const_iterator &it, string::const_iterator &it_end, char delim );
    void addGenotypeRow( int rIdx, const char *p_begin, const char *p_end, char delim );

    ushort encodeGenotype( const string &gt );
    const char *decodeGenotype( ushort encoded_gt );
    bool isGenotypeHomozygous( ushort encoded_gt );

    void selectMarker( uint rIdx );
    void getGenotypeDistribution( uint rIdx, GenotypeDistribution &dist );
    void getCaseControlGenotypeDistribution( uint rIdx, CaseControlSet &ccs, CaseControlGenotypeDistribution &ccgd );
    void getCaseControlGenotypeDistribution( uint rIdx, CaseControlGenotypeDistribution &ccgd );
    void getCaseControlGenotypeDistribution( uint rIdx, CaseControlGenotypeDistribution &ccgd, marginal_information &m );

    void selectMarkerPair( uint maIdx, uint mbIdx );
    void selectCaseControl( CaseControlSet &ccs );

    void getContingencyTable( uint rIdx1, uint rIdx2, ContingencyTable &ct );
    void getContingencyTable( uint rIdx1, uint rIdx2, ushort *column_set, ContingencyTable &ct );
    void getCaseControlContingencyTable( uint rIdx1, uint rIdx2, CaseControlSet &ccs, CaseControlContingencyTable &ccct );
    void getCaseControlContingencyTable( uint rIdx1, uint rIdx2, CaseControlContingencyTable &ccct );
    void getCaseControlContingencyTable( uint rIdx1, uint rIdx2, const marginal_information &m1, const marginal_information &m2, CaseControlContingencyTable &ccct );

    gwasdev_store *handle() { flush(); return store; }        // for callers that want the batch entry points of gwasdev.h

    virtual ~DeviceGenotypeTable();
protected:
    void initialize();

private:
    void flush();                                   // rows packed on the host since the last call go to the device
    void invalidateCaches();
    void ensureStreamMasks( CaseControlSet &ccs );
    const uint *rowBlock( int mode, uint rIdx );    // 4 (mode 0) or 8 counts of row rIdx through overload `mode`
    const uint *pairBlock( int mode, uint i, uint j );   // 32 counts: case table, control table

    gwasdev_store *store;
    uint plane_blocks;
    std::vector<ushort> pending;
    int pending_first, pending_count;
    std::vector<ushort> cell_row, fly_masks;
    char call_buf[3];
    char gt_text[17][3];

    // blocks served to the per-item overloads
    int row_mode; uint row_first, row_count; size_t row_block;
    std::vector<uint> row_counts;
    std::vector<marginal_information> row_margins;
    int pair_mode; uint pair_i, pair_j0, pair_count; size_t pair_block;
    std::vector<uint> pair_pi, pair_pj, pair_tables;
};

}
}

#endif // DEVICE_GENOTYPE_TABLE_H
