// See device_geno_table.h. Error convention of the reference on this path: assert -> abort.
#include "device_geno_table.h"

#include <cassert>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace libgwaspp {
namespace genetics {

#define GW_MUST(call)                                                                         \
    do {                                                                                      \
        const int rc_ = (call);                                                               \
        if (rc_ != GWASDEV_OK) {                                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, gwasdev_last_error());        \
            assert(false && "gwasdev call failed");                                           \
            abort();                                                                          \
        }                                                                                     \
    } while (0)

static const size_t FLUSH_ROWS = 8192;

// ---- CaseControlSet ------------------------------------------------------------------------------
CaseControlSet::CaseControlSet(int n) : max_index((uint)n - 1), case_count(0), ctrl_count(0), rev(1) {
    const uint P = gwasdev_plane_blocks((uint)n);
    stream_case_set.assign(P, 0);
    stream_control_set.assign(P, 0);
}
void CaseControlSet::reset() {
    std::fill(stream_case_set.begin(), stream_case_set.end(), 0);
    std::fill(stream_control_set.begin(), stream_control_set.end(), 0);
    case_count = ctrl_count = 0;
    ++rev;
}
void CaseControlSet::setCases(const std::set<int> &idx) {
    std::fill(stream_case_set.begin(), stream_case_set.end(), 0);
    for (int i : idx) { assert(i >= 0 && i <= (int)max_index); stream_case_set[i >> 4] |= (ushort)(1u << (i & 15)); }
    case_count = (uint)idx.size();
    ++rev;
}
void CaseControlSet::setControls(const std::set<int> &idx) {
    std::fill(stream_control_set.begin(), stream_control_set.end(), 0);
    for (int i : idx) { assert(i >= 0 && i <= (int)max_index); stream_control_set[i >> 4] |= (ushort)(1u << (i & 15)); }
    ctrl_count = (uint)idx.size();
    ++rev;
}
void CaseControlSet::setAllAsCases() {
    reset();
    for (uint i = 0; i <= max_index; ++i) stream_case_set[i >> 4] |= (ushort)(1u << (i & 15));
    case_count = max_index + 1;
}
void CaseControlSet::setAllAsControls() {
    reset();
    for (uint i = 0; i <= max_index; ++i) stream_control_set[i >> 4] |= (ushort)(1u << (i & 15));
    ctrl_count = max_index + 1;
}

// ---- DeviceGenoTable -----------------------------------------------------------------------------
DeviceGenoTable::DeviceGenoTable(int n_markers, int n_individuals, int device)
    : max_row(n_markers), max_column(n_individuals), store(nullptr), pending_first(0), pending_count(0),
      selected_rev(0), selected_set(nullptr), fly_rev(0), fly_set(nullptr), device0(device), n_devices(1), replica_rev(0) {
    GW_MUST(gwasdev_create((uint64_t)n_markers, (uint32_t)n_individuals, device, &store));
    plane_blocks = gwasdev_plane_blocks((uint32_t)n_individuals);
    cell_row.resize(2 * plane_blocks + 1);
    call_buf[0] = call_buf[1] = call_buf[2] = 0;
}

// table sized from and loaded with a transposed-PLINK genotype file in one pass (plain) or two (.gz), parsed on the device
DeviceGenoTable::DeviceGenoTable(const std::string &tped_path, int device)
    : max_row(0), max_column(0), store(nullptr), pending_first(0), pending_count(0), selected_rev(0), selected_set(nullptr),
      fly_rev(0), fly_set(nullptr), device0(device), n_devices(1), replica_rev(0) {
    uint64_t rows = 0;
    uint32_t cols = 0;
    GW_MUST(gwasdev_create_from_tped(tped_path.c_str(), device, &store, &rows, &cols));
    max_row = (int)rows;
    max_column = (int)cols;
    plane_blocks = gwasdev_plane_blocks(cols);
    cell_row.resize(2 * plane_blocks + 1);
    call_buf[0] = call_buf[1] = call_buf[2] = 0;
}

DeviceGenoTable::~DeviceGenoTable() { dropReplicas(); gwasdev_destroy(store); }

void DeviceGenoTable::dropReplicas() {
    for (gwasdev_store *r : replicas) gwasdev_destroy(r);
    replicas.clear();
    replica_rev = 0;
}

void DeviceGenoTable::useDevices(int n) {
    assert(n >= 1 && device0 + n <= gwasdev_device_count());
    if (n != n_devices) dropReplicas();
    n_devices = n;
}

// replicas of the table on devices device0 + 1 .. device0 + n_devices - 1, carrying the current selection
void DeviceGenoTable::syncReplicas() {
    if (n_devices <= 1) return;
    assert(selected_set != nullptr && selected_rev != 0);
    if (replicas.empty()) {
        for (int d = 1; d < n_devices; ++d) {
            gwasdev_store *r = nullptr;
            GW_MUST(gwasdev_replicate(store, device0 + d, &r));      // table + selection, device to device
            replicas.push_back(r);
        }
    } else if (replica_rev != selected_rev) {
        CaseControlSet &ccs = *const_cast<CaseControlSet *>(selected_set);
        for (gwasdev_store *r : replicas) GW_MUST(gwasdev_select_case_control(r, ccs.stream_case_begin(), ccs.stream_control_begin()));
    }
    replica_rev = selected_rev;
}

void DeviceGenoTable::flush() {
    if (pending_count == 0) return;
    GW_MUST(gwasdev_put_rows(store, (uint64_t)pending_first, (uint64_t)pending_count, pending.data()));
    pending_count = 0;
    selected_rev = 0;
    dropReplicas();
}

int DeviceGenoTable::loadTransposedPlink(const std::string &tped_path, int first_row) {
    flush();
    uint64_t rows = 0;
    GW_MUST(gwasdev_load_tped(store, tped_path.c_str(), (uint64_t)first_row, &rows));
    selected_rev = 0;
    dropReplicas();
    return (int)rows;
}

int DeviceGenoTable::loadBed(const std::string &bed_path, const std::vector<unsigned char> &alleles, int first_row) {
    flush();
    uint64_t rows = 0;
    GW_MUST(gwasdev_load_bed(store, bed_path.c_str(), alleles.empty() ? nullptr : alleles.data(), (uint64_t)first_row, &rows));
    selected_rev = 0;
    dropReplicas();
    return (int)rows;
}

void DeviceGenoTable::addGenotypeRow(int rIdx, const char *p_begin, const char *p_end, char /*delim*/) {
    if (p_begin >= p_end) return;
    assert(rIdx >= 0 && rIdx < max_row);
    const size_t row_len = 2 * (size_t)plane_blocks + 1;
    if (pending_count > 0 && (rIdx != pending_first + pending_count || (size_t)pending_count >= FLUSH_ROWS)) flush();
    if (pending_count == 0) pending_first = rIdx;
    if (pending.size() < (size_t)(pending_count + 1) * row_len) pending.resize((size_t)(pending_count + 1) * row_len);
    GW_MUST(gwasdev_pack_row_text(p_begin, (size_t)(p_end - p_begin), (uint32_t)max_column, pending.data() + (size_t)pending_count * row_len));
    ++pending_count;
}

void DeviceGenoTable::addGenotypeRow(int rIdx, std::string::const_iterator &it, std::string::const_iterator &it_end, char delim) {
    if (it >= it_end) return;
    addGenotypeRow(rIdx, &*it, &*it + (it_end - it), delim);
}

// Single-cell update. Re-labels the whole row with the row loader's first-seen rule (the reference's own
// single-cell path passes two arguments of its header state machine in swapped order,
// compressed_genotype_table5.cpp:173; rows should be loaded through addGenotypeRow).
void DeviceGenoTable::addGenotype(int rIdx, int cIdx, const std::string &gt) {
    assert(gt.length() == 2 && rIdx >= 0 && rIdx < max_row && cIdx >= 0 && cIdx < max_column);
    flush();
    std::string line((size_t)max_column * 3, '\t');
    for (int c = 0; c < max_column; ++c) {
        const char *call = c == cIdx ? gt.c_str() : getCallAt((uint)rIdx, (uint)c);
        line[3 * c] = call[0];
        line[3 * c + 1] = call[1];
    }
    GW_MUST(gwasdev_pack_row_text(line.data(), line.size() - 1, (uint32_t)max_column, cell_row.data()));
    GW_MUST(gwasdev_put_rows(store, (uint64_t)rIdx, 1, cell_row.data()));
    selected_rev = 0;
    dropReplicas();
}

static int allele_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }

ushort DeviceGenoTable::encodeGenotype(const std::string &gt) {
    assert(gt.length() == 2);
    const int a = allele_index(gt[0]), b = allele_index(gt[1]);
    return (a < 4 && b < 4) ? (ushort)(4 * a + b) : (ushort)0xFFFF;
}

const char *DeviceGenoTable::decodeGenotype(ushort enc) {
    static const char table[17][3] = {"AA", "AC", "AG", "AT", "CA", "CC", "CG", "CT", "GA", "GC", "GG", "GT",
                                      "TA", "TC", "TG", "TT", "00"};
    return enc < 16 ? table[enc] : table[16];
}

bool DeviceGenoTable::isGenotypeHomozygous(ushort enc) { return enc == 0 || enc == 5 || enc == 10 || enc == 15; }

DataBlock DeviceGenoTable::operator()(int r, int c) {
    flush();
    GW_MUST(gwasdev_call_at(store, (uint64_t)r, (uint32_t)c, call_buf));
    return encodeGenotype(std::string(call_buf, 2));
}

void DeviceGenoTable::selectCaseControl(CaseControlSet &ccs) {
    flush();
    GW_MUST(gwasdev_select_case_control(store, ccs.stream_case_begin(), ccs.stream_control_begin()));
    selected_rev = fly_rev = ccs.revision();      // a selection also sets the on-the-fly masks (to its own)
    selected_set = fly_set = &ccs;
}

// Masks of the mask-on-the-fly overloads. As in the reference (compressed_genotype_table5.cpp:609-657, :806-895) these never
// touch the pre-selected store: a caller may select with one set, probe on the fly with another and go on using the
// pre-selected overloads.
void DeviceGenoTable::ensureMasks(CaseControlSet &ccs) {
    flush();
    if (fly_set == &ccs && fly_rev == ccs.revision()) return;
    GW_MUST(gwasdev_set_stream_masks(store, ccs.stream_case_begin(), ccs.stream_control_begin()));
    fly_set = &ccs;
    fly_rev = ccs.revision();
}

void DeviceGenoTable::selectMarker(uint) { assert(false); }            // as in the reference's bit-plane tables
void DeviceGenoTable::selectMarkerPair(uint, uint) { assert(false); }

void DeviceGenoTable::getGenotypeDistribution(uint rIdx, GenotypeDistribution &dist) {
    flush();
    frequency_table ft;
    GW_MUST(gwasdev_counts(store, rIdx, rIdx + 1, 0, ft.freq));
    dist.setDistribution(ft);
    dist.setCurrentIndex(rIdx);
}

void DeviceGenoTable::getCaseControlGenotypeDistribution(uint rIdx, CaseControlSet &ccs, CaseControlGenotypeDistribution &ccgd) {
    ensureMasks(ccs);
    uint c[8];
    GW_MUST(gwasdev_counts(store, rIdx, rIdx + 1, 1, c));
    frequency_table a, b;
    memcpy(a.freq, c, 16); memcpy(b.freq, c + 4, 16);
    ccgd.setCaseDistribution(a); ccgd.setControlDistribution(b); ccgd.setCurrentIndex(rIdx);
}

void DeviceGenoTable::getCaseControlGenotypeDistribution(uint rIdx, CaseControlGenotypeDistribution &ccgd) {
    flush();
    uint c[8];
    GW_MUST(gwasdev_counts(store, rIdx, rIdx + 1, 2, c));
    frequency_table a, b;
    memcpy(a.freq, c, 16); memcpy(b.freq, c + 4, 16);
    ccgd.setCaseDistribution(a); ccgd.setControlDistribution(b); ccgd.setCurrentIndex(rIdx);
}

void DeviceGenoTable::getCaseControlGenotypeDistribution(uint rIdx, CaseControlGenotypeDistribution &ccgd, marginal_information &m) {
    flush();
    uint c[8];
    GW_MUST(gwasdev_marginal_scan(store, rIdx, rIdx + 1, c, reinterpret_cast<gwasdev_marginal_information *>(&m), nullptr, 0));
    frequency_table a, b;
    memcpy(a.freq, c, 16); memcpy(b.freq, c + 4, 16);
    ccgd.setCaseDistribution(a); ccgd.setControlDistribution(b); ccgd.setCurrentIndex(rIdx);
}

static void one_table(gwasdev_store *store, uint i, uint j, int mode, CONTIN_TABLE_T &ca, CONTIN_TABLE_T &co) {
    uint out[32];
    GW_MUST(gwasdev_pair_tables(store, 1, &i, &j, mode, out));
    memcpy(ca.contin, out, 64);
    memcpy(co.contin, out + 16, 64);
}

void DeviceGenoTable::getContingencyTable(uint rIdx1, uint rIdx2, ContingencyTable &ct) {
    flush();
    CONTIN_TABLE_T a, b;
    one_table(store, rIdx1, rIdx2, 0, a, b);
    ct.setMarkerAIndex(rIdx1); ct.setMarkerBIndex(rIdx2);
    ct.setContingency(a);
}

void DeviceGenoTable::getContingencyTable(uint, uint, ushort *, ContingencyTable &) { assert(false); }   // reference: assert(false)

void DeviceGenoTable::getCaseControlContingencyTable(uint rIdx1, uint rIdx2, CaseControlSet &ccs, CaseControlContingencyTable &ccct) {
    ensureMasks(ccs);
    CONTIN_TABLE_T a, b;
    one_table(store, rIdx1, rIdx2, 1, a, b);
    ccct.setMarkerAIndex(rIdx1); ccct.setMarkerBIndex(rIdx2);
    ccct.updateContingencyTables(a, b);
}

void DeviceGenoTable::getCaseControlContingencyTable(uint rIdx1, uint rIdx2, CaseControlContingencyTable &ccct) {
    flush();
    CONTIN_TABLE_T a, b;
    one_table(store, rIdx1, rIdx2, 2, a, b);
    ccct.setMarkerAIndex(rIdx1); ccct.setMarkerBIndex(rIdx2);
    ccct.updateContingencyTables(a, b);
}

// The margins are recomputed (and cached) on the device from the same compacted rows, so m1/m2 only
// select this overload; they are not uploaded.
void DeviceGenoTable::getCaseControlContingencyTable(uint rIdx1, uint rIdx2, const marginal_information &, const marginal_information &,
                                                     CaseControlContingencyTable &ccct) {
    flush();
    CONTIN_TABLE_T a, b;
    one_table(store, rIdx1, rIdx2, 3, a, b);
    ccct.setMarkerAIndex(rIdx1); ccct.setMarkerBIndex(rIdx2);
    ccct.updateContingencyTables(a, b);
}

// ---- batch entry points ----------------------------------------------------------------------------
void DeviceGenoTable::computeMargins(std::vector<marginal_information> &out) {
    flush();
    out.resize((size_t)max_row);
    GW_MUST(gwasdev_marginal_scan(store, 0, (uint64_t)max_row, nullptr, reinterpret_cast<gwasdev_marginal_information *>(out.data()), nullptr, 0));
}

void DeviceGenoTable::scanCaseControl(std::vector<frequency_table> &cases, std::vector<frequency_table> &controls,
                                      std::vector<gwasdev_snp_stats> *stats) {
    flush();
    std::vector<uint> c((size_t)max_row * 8);
    if (stats) stats->resize((size_t)max_row);
    GW_MUST(gwasdev_marginal_scan(store, 0, (uint64_t)max_row, c.data(), nullptr, stats ? stats->data() : nullptr, 0));
    cases.resize((size_t)max_row); controls.resize((size_t)max_row);
    for (int r = 0; r < max_row; ++r) { memcpy(cases[r].freq, &c[8 * (size_t)r], 16); memcpy(controls[r].freq, &c[8 * (size_t)r + 4], 16); }
}

void DeviceGenoTable::screenPairs(double threshold, std::vector<gwasdev_hit> &hits, gwasdev_pair_stats *stats, uint shard, uint n_shards) {
    flush();
    uint64_t n = 0;
    if (n_devices > 1) {     // all shards at once, one per device, inside the library
        assert(shard == 0 && n_shards == 1);
        syncReplicas();
        std::vector<gwasdev_store *> all(1, store);
        all.insert(all.end(), replicas.begin(), replicas.end());
        std::vector<gwasdev_pair_stats> st(all.size());
        if (hits.size() < (1u << 22)) hits.resize(1u << 22);
        int rc = gwasdev_pairwise_scan_multi(all.data(), (uint32_t)all.size(), threshold, 0, hits.data(), hits.size(), &n, st.data(), 0);
        if (rc == GWASDEV_EOVERFLOW) {
            hits.resize((size_t)n);
            rc = gwasdev_pairwise_scan_multi(all.data(), (uint32_t)all.size(), threshold, 0, hits.data(), hits.size(), &n, st.data(), 0);
        }
        GW_MUST(rc);
        hits.resize((size_t)n);
        if (stats) {         // totals over the shards; times are the slowest shard's
            *stats = st[0];
            for (size_t d = 1; d < st.size(); ++d) {
                stats->pairs_tested += st[d].pairs_tested; stats->candidates += st[d].candidates; stats->word_cells += st[d].word_cells;
                stats->tiles += st[d].tiles; stats->tiles_nine_cell += st[d].tiles_nine_cell;
                stats->screen_ms = std::max(stats->screen_ms, st[d].screen_ms); stats->total_ms = std::max(stats->total_ms, st[d].total_ms);
            }
            stats->hits = n;
        }
        return;
    }
    if (hits.size() < 1024) hits.resize(1024);
    int rc = gwasdev_pairwise_scan(store, threshold, shard, n_shards, hits.data(), hits.size(), &n, stats, 0);
    if (rc == GWASDEV_EOVERFLOW) {
        hits.resize((size_t)n);
        rc = gwasdev_pairwise_scan(store, threshold, shard, n_shards, hits.data(), hits.size(), &n, stats, 0);
    }
    GW_MUST(rc);
    hits.resize((size_t)n);
}

void DeviceGenoTable::gtestPairs(const std::vector<gwasdev_hit> &hits, std::vector<double> &stat, std::vector<double> &z) {
    flush();
    std::vector<uint> pi(hits.size()), pj(hits.size());
    for (size_t k = 0; k < hits.size(); ++k) { pi[k] = hits[k].i; pj[k] = hits[k].j; }
    stat.resize(hits.size()); z.resize(hits.size());
    if (hits.empty()) return;
    GW_MUST(gwasdev_gtest(store, hits.size(), pi.data(), pj.data(), stat.data(), z.data()));
}

}  // namespace genetics
}  // namespace libgwaspp
