// See device_genotype_table.h. Error convention of the reference on this path: assert -> abort, no exceptions, no codes.
#include "device_genotype_table.h"

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace libgwaspp {
namespace genetics {

#define DEV_MUST(call)                                                                        \
    do {                                                                                      \
        const int rc_ = (call);                                                               \
        if (rc_ != GWASDEV_OK) {                                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, gwasdev_last_error());        \
            assert(false && "gwasdev call failed");                                           \
            abort();                                                                          \
        }                                                                                     \
    } while (0)

// items fetched per batched call: 64 for an isolated request, doubling while requests continue where the last block ended
static const size_t FLUSH_ROWS = 8192, ROW_BLOCK_MAX = 8192, PAIR_BLOCK_MAX = 4096, BLOCK_MIN = 64;

DeviceGenotypeTable::DeviceGenotypeTable( util::indexer *markers, util::indexer *individs, int device )
    : GenoTable( markers, individs ), store( NULL ), pending_first( 0 ), pending_count( 0 ), row_mode( -1 ), row_first( 0 ), row_count( 0 ),
      row_block( 64 ), pair_mode( -1 ), pair_i( 0 ), pair_j0( 0 ), pair_count( 0 ), pair_block( 64 ) {
    DEV_MUST( gwasdev_create( (uint64_t) max_row, (uint32_t) max_column, device, &store ) );
    initialize();
}

// What CompressedGenotypeTable5::initialize (compressed_genotype_table5.cpp:34-153) sets up on the host, minus the table
// itself: the geometry fields GenoTable declares, the 16 + 1 genotype codes and their spellings.
void DeviceGenotypeTable::initialize() {
    alphabet_size = 4;
    possible_genotypes_size = 17;
    bits_per_data = 1;
    data_per_block = 16;
    plane_blocks = gwasdev_plane_blocks( (uint32_t) max_column );
    blocks_per_row = 2 * plane_blocks + 1;
    total_block_count = blocks_per_row * max_row;
    bytes_per_row = blocks_per_row * sizeof( DataBlock );
    data_size = total_block_count * sizeof( DataBlock );
    beg = new ushort[ possible_genotypes_size ];
    for( uint k = 0; k < 16; ++k ) beg[k] = (ushort) k;
    beg[16] = 0xFFFF;
    end = beg + possible_genotypes_size;
    static const char acgt[5] = "ACGT";
    memset( transformations, 4, 256 );
    for( int k = 0; k < 4; ++k ) transformations[ (byte) acgt[k] ] = (byte) k;
    for( int k = 0; k < 16; ++k ) { gt_text[k][0] = acgt[k >> 2]; gt_text[k][1] = acgt[k & 3]; gt_text[k][2] = 0; }
    gt_text[16][0] = gt_text[16][1] = '0'; gt_text[16][2] = 0;
    cell_row.resize( 2 * plane_blocks + 1 );
    call_buf[0] = call_buf[1] = call_buf[2] = 0;
}

DeviceGenotypeTable::~DeviceGenotypeTable() { gwasdev_destroy( store ); }

void DeviceGenotypeTable::invalidateCaches() { row_mode = -1; pair_mode = -1; }

void DeviceGenotypeTable::flush() {
    if( pending_count == 0 ) return;
    DEV_MUST( gwasdev_put_rows( store, (uint64_t) pending_first, (uint64_t) pending_count, pending.data() ) );
    pending_count = 0;
    invalidateCaches();
}

void DeviceGenotypeTable::addGenotypeRow( int rIdx, const char *p_begin, const char *p_end, char ) {
    if( p_begin >= p_end ) return;
    assert( rIdx >= 0 && rIdx < max_row );
    const size_t row_len = 2 * (size_t) plane_blocks + 1;
    if( pending_count > 0 && ( rIdx != pending_first + pending_count || (size_t) pending_count >= FLUSH_ROWS ) ) flush();
    if( pending_count == 0 ) pending_first = rIdx;
    if( pending.size() < (size_t)( pending_count + 1 ) * row_len ) pending.resize( (size_t)( pending_count + 1 ) * row_len );
    DEV_MUST( gwasdev_pack_row_text( p_begin, (size_t)( p_end - p_begin ), (uint32_t) max_column, pending.data() + (size_t) pending_count * row_len ) );
    ++pending_count;
}

void DeviceGenotypeTable::addGenotypeRow( int rIdx, string::const_iterator &it, string::const_iterator &it_end, char delim ) {
    if( it >= it_end ) return;
    addGenotypeRow( rIdx, &*it, &*it + ( it_end - it ), delim );
}

// Single-cell update: the row is re-labelled with the row loader's first-seen rule (the reference's own single-cell path
// hands two arguments of its header state machine over in swapped order, compressed_genotype_table5.cpp:173).
void DeviceGenotypeTable::addGenotype( int rIdx, int cIdx, const string &gt ) {
    assert( gt.length() == 2 && rIdx >= 0 && rIdx < max_row && cIdx >= 0 && cIdx < max_column );
    flush();
    string line( (size_t) max_column * 3, '\t' );
    for( int c = 0; c < max_column; ++c ) {
        const char *call = c == cIdx ? gt.c_str() : getCallAt( (uint) rIdx, (uint) c );
        line[3 * c] = call[0];
        line[3 * c + 1] = call[1];
    }
    DEV_MUST( gwasdev_pack_row_text( line.data(), line.size() - 1, (uint32_t) max_column, cell_row.data() ) );
    DEV_MUST( gwasdev_put_rows( store, (uint64_t) rIdx, 1, cell_row.data() ) );
    invalidateCaches();
}

ushort DeviceGenotypeTable::encodeGenotype( const string &gt ) {
    assert( gt.length() == 2 );
    const int a = transformations[ (byte) gt[0] ], b = transformations[ (byte) gt[1] ];
    return ( a < 4 && b < 4 ) ? (ushort)( 4 * a + b ) : (ushort) 0xFFFF;
}

const char *DeviceGenotypeTable::decodeGenotype( ushort enc ) { return enc < 16 ? gt_text[enc] : gt_text[16]; }

bool DeviceGenotypeTable::isGenotypeHomozygous( ushort enc ) { return enc == 0 || enc == 5 || enc == 10 || enc == 15; }

DataBlock DeviceGenotypeTable::operator()( int r, int c ) {
    flush();
    DEV_MUST( gwasdev_call_at( store, (uint64_t) r, (uint32_t) c, call_buf ) );
    return (DataBlock) encodeGenotype( string( call_buf, 2 ) );
}

void DeviceGenotypeTable::selectMarker( uint ) { assert( false ); }              // as the reference's bit-plane tables
void DeviceGenotypeTable::selectMarkerPair( uint, uint ) { assert( false ); }

void DeviceGenotypeTable::selectCaseControl( CaseControlSet &ccs ) {
    flush();
    DEV_MUST( gwasdev_select_case_control( store, ccs.stream_case_begin(), ccs.stream_control_begin() ) );
    nCaseCount = ccs.getCaseCount(); nControlCount = ccs.getControlCount(); nIndivids = nCaseCount + nControlCount;
    fly_masks.assign( ccs.stream_case_begin(), ccs.stream_case_begin() + plane_blocks );
    fly_masks.insert( fly_masks.end(), ccs.stream_control_begin(), ccs.stream_control_begin() + plane_blocks );
    invalidateCaches();
}

// The mask-on-the-fly overloads get the set with every call; its masks go to the device when they differ from the ones
// there (2 x P blocks compared on the host). The pre-selected store is not touched, as in the reference (:609-657, :806-895).
void DeviceGenotypeTable::ensureStreamMasks( CaseControlSet &ccs ) {
    flush();
    const size_t P = plane_blocks;
    if( fly_masks.size() == 2 * P && memcmp( fly_masks.data(), ccs.stream_case_begin(), 2 * P ) == 0 &&
        memcmp( fly_masks.data() + P, ccs.stream_control_begin(), 2 * P ) == 0 ) return;
    DEV_MUST( gwasdev_set_stream_masks( store, ccs.stream_case_begin(), ccs.stream_control_begin() ) );
    fly_masks.assign( ccs.stream_case_begin(), ccs.stream_case_begin() + P );
    fly_masks.insert( fly_masks.end(), ccs.stream_control_begin(), ccs.stream_control_begin() + P );
    if( row_mode == 1 ) row_mode = -1;
    if( pair_mode == 1 ) pair_mode = -1;
}

// mode 0: whole cohort (4 counts); 1: mask-on-the-fly; 2: pre-selected; 3: pre-selected + marginal_information
const uint *DeviceGenotypeTable::rowBlock( int mode, uint rIdx ) {
    assert( (int) rIdx < max_row );
    const uint per = mode == 0 ? 4 : 8;
    if( row_mode != mode || rIdx < row_first || rIdx >= row_first + row_count ) {
        flush();
        row_block = ( row_mode == mode && rIdx == row_first + row_count ) ? std::min( ROW_BLOCK_MAX, 2 * row_block ) : BLOCK_MIN;
        row_first = rIdx;
        row_count = (uint) std::min<size_t>( row_block, (size_t) max_row - rIdx );
        row_counts.resize( (size_t) row_count * per );
        if( mode == 3 ) {
            row_margins.resize( row_count );
            DEV_MUST( gwasdev_marginal_scan( store, row_first, row_first + row_count, row_counts.data(),
                                             reinterpret_cast<gwasdev_marginal_information *>( row_margins.data() ), NULL, 0 ) );
        } else DEV_MUST( gwasdev_counts( store, row_first, row_first + row_count, mode, row_counts.data() ) );
        row_mode = mode;
    }
    return row_counts.data() + (size_t)( rIdx - row_first ) * per;
}

void DeviceGenotypeTable::getGenotypeDistribution( uint rIdx, GenotypeDistribution &dist ) {
    frequency_table ft;
    memcpy( ft.freq, rowBlock( 0, rIdx ), 16 );
    dist.setDistribution( ft );
}

static void fill_ccgd( const uint *c, CaseControlGenotypeDistribution &ccgd ) {
    frequency_table a, b;
    memcpy( a.freq, c, 16 ); memcpy( b.freq, c + 4, 16 );
    ccgd.setCaseDistribution( a ); ccgd.setControlDistribution( b );
}

void DeviceGenotypeTable::getCaseControlGenotypeDistribution( uint rIdx, CaseControlSet &ccs, CaseControlGenotypeDistribution &ccgd ) {
    ensureStreamMasks( ccs );
    fill_ccgd( rowBlock( 1, rIdx ), ccgd );
}

void DeviceGenotypeTable::getCaseControlGenotypeDistribution( uint rIdx, CaseControlGenotypeDistribution &ccgd ) {
    fill_ccgd( rowBlock( 2, rIdx ), ccgd );
}

void DeviceGenotypeTable::getCaseControlGenotypeDistribution( uint rIdx, CaseControlGenotypeDistribution &ccgd, marginal_information &m ) {
    fill_ccgd( rowBlock( 3, rIdx ), ccgd );
    m = row_margins[ rIdx - row_first ];
}

// mode as gwasdev_pair_tables: 0 un-stratified, 1 mask-on-the-fly, 2 pre-selected, 3 margins overload
const uint *DeviceGenotypeTable::pairBlock( int mode, uint i, uint j ) {
    assert( (int) i < max_row && (int) j < max_row );
    if( pair_mode != mode || i != pair_i || j < pair_j0 || j >= pair_j0 + pair_count ) {
        flush();
        pair_block = ( pair_mode == mode && i == pair_i && j == pair_j0 + pair_count ) ? std::min( PAIR_BLOCK_MAX, 2 * pair_block ) : BLOCK_MIN;
        pair_i = i; pair_j0 = j;
        pair_count = (uint) std::min<size_t>( pair_block, (size_t) max_row - j );
        pair_pi.assign( pair_count, i );
        pair_pj.resize( pair_count );
        for( uint q = 0; q < pair_count; ++q ) pair_pj[q] = j + q;
        pair_tables.resize( (size_t) pair_count * 32 );
        DEV_MUST( gwasdev_pair_tables( store, pair_count, pair_pi.data(), pair_pj.data(), mode, pair_tables.data() ) );
        pair_mode = mode;
    }
    return pair_tables.data() + (size_t)( j - pair_j0 ) * 32;
}

void DeviceGenotypeTable::getContingencyTable( uint rIdx1, uint rIdx2, ContingencyTable &ct ) {
    CONTIN_TABLE_T a;
    memcpy( a.contin, pairBlock( 0, rIdx1, rIdx2 ), 64 );
    ct.setMarkerAIndex( rIdx1 ); ct.setMarkerBIndex( rIdx2 );
    ct.setContingency( a );
}

void DeviceGenotypeTable::getContingencyTable( uint, uint, ushort *, ContingencyTable & ) { assert( false ); }   // reference: assert(false)

static void fill_ccct( const uint *t, uint i, uint j, CaseControlContingencyTable &ccct ) {
    CONTIN_TABLE_T a, b;
    memcpy( a.contin, t, 64 ); memcpy( b.contin, t + 16, 64 );
    ccct.setMarkerAIndex( i ); ccct.setMarkerBIndex( j );
    ccct.updateContingencyTables( a, b );
}

void DeviceGenotypeTable::getCaseControlContingencyTable( uint rIdx1, uint rIdx2, CaseControlSet &ccs, CaseControlContingencyTable &ccct ) {
    ensureStreamMasks( ccs );
    fill_ccct( pairBlock( 1, rIdx1, rIdx2 ), rIdx1, rIdx2, ccct );
}

void DeviceGenotypeTable::getCaseControlContingencyTable( uint rIdx1, uint rIdx2, CaseControlContingencyTable &ccct ) {
    fill_ccct( pairBlock( 2, rIdx1, rIdx2 ), rIdx1, rIdx2, ccct );
}

// m1 / m2 only select this overload: the device derives the same margins from the same pre-selected rows (and caches them).
void DeviceGenotypeTable::getCaseControlContingencyTable( uint rIdx1, uint rIdx2, const marginal_information &, const marginal_information &,
                                                          CaseControlContingencyTable &ccct ) {
    fill_ccct( pairBlock( 3, rIdx1, rIdx2 ), rIdx1, rIdx2, ccct );
}

}
}
