// f1: genotype files -> device store without a host parser.
//
// The reference parses transposed-PLINK text on one host core, one stream extraction per line
// (genetics/individual/tped_genotype_file.cpp:127-190: trim, skip four marker fields, alleles 1234 -> ACGT, every other
// character taken as an allele) and packs each row bit by bit with the first-seen label state machine
// (genetics/genotype/compressed_genotype_table5.cpp:277-365); at configs[0] that is 0.42 s of a 0.7 s run. Here the
// host only moves bytes: file -> pinned buffer -> HBM. On the device
//   1. newline index: two passes over the text (count per 4 KiB block, exclusive scan, ordered write),
//   2. line spans: trim, drop empty lines, row rank by exclusive scan,
//   3. one warp per line: 32 genotypes per step, first-seen labels resolved in column order with ballots (at most
//      three new labels per row), bit-planes assembled by __ballot_sync and written 128 bytes at a time
// and the same row assembler takes PLINK .bed rows (SNP-major, 2 bits per genotype), which the reference cannot read.
// Files are read through zlib, which also reads plain files; a .gz is simply opened twice (dims pass, load pass), the
// rewind the reference's gzstream cannot do (individual_genotype_file.cpp:94, SURVEY.md inventory row 9).
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace gwasdev {

constexpr int NL_THREADS = 256;
constexpr uint32_t NL_BLOCK = NL_THREADS * 16;   // bytes per block of the newline passes

enum IngestError : int { ING_OK = 0, ING_THIRD_SPELLING = 1, ING_TOO_MANY_ROWS = 2, ING_SHORT_HEADER = 3, ING_TOO_MANY_LINES = 4 };

struct IngestCounters {
    unsigned long long n_lines;   // '\n'-terminated lines in the chunk
    unsigned long long n_rows;    // non-empty lines = table rows written
    unsigned long long cursor;    // table row of the next non-empty line (advanced after every chunk)
    unsigned long long err_row;
    uint32_t err_col;
    int err;
};

struct Ingest {
    IngestCounters *d_cnt = nullptr, *h_cnt = nullptr;
    char *h_pin[2] = {nullptr, nullptr};
    char *d_text[2] = {nullptr, nullptr};
    size_t cap_pin = 0, cap_text = 0;
    uint32_t *d_block_cnt = nullptr, *d_block_off = nullptr, *d_nl = nullptr, *d_flag = nullptr, *d_rank = nullptr;
    uint2 *d_span = nullptr;
    size_t cap_blocks = 0, cap_lines = 0;
    void *d_tmp = nullptr;
    size_t cap_tmp = 0;
    uint8_t *d_alleles = nullptr;
    size_t cap_alleles = 0;
    cudaEvent_t done[2] = {nullptr, nullptr};
};

__device__ __forceinline__ uint32_t newline_mask16(const uint4 v) {   // bit k set when byte k of the 16 is '\n'
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t eq = __vcmpeq4(w[q], 0x0A0A0A0Au) & 0x01010101u;            // 0x01 per matching byte
        m |= (((eq * 0x01020408u) >> 24) & 0xFu) << (4 * q);                       // byte k's flag -> bit 24 + k
    }
    return m;
}

__global__ void __launch_bounds__(NL_THREADS) newline_count_kernel(const uint4 *__restrict__ text, uint32_t *__restrict__ block_cnt) {
    const uint32_t n = __popc(newline_mask16(text[(size_t)blockIdx.x * NL_THREADS + threadIdx.x]));
    typedef cub::BlockReduce<uint32_t, NL_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const uint32_t total = Reduce(tmp).Sum(n);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(NL_THREADS) newline_write_kernel(const uint4 *__restrict__ text, const uint32_t *__restrict__ block_cnt,
                                                                   const uint32_t *__restrict__ block_off, uint32_t *__restrict__ nl,
                                                                   uint32_t cap_lines, IngestCounters *cnt) {
    uint32_t m = newline_mask16(text[(size_t)blockIdx.x * NL_THREADS + threadIdx.x]);
    typedef cub::BlockScan<uint32_t, NL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    uint32_t before;
    Scan(tmp).ExclusiveSum((uint32_t)__popc(m), before);
    uint32_t o = block_off[blockIdx.x] + before;
    const uint32_t pos0 = (blockIdx.x * NL_THREADS + threadIdx.x) * 16u;
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        if (o < cap_lines) nl[o] = pos0 + k;
        ++o;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const unsigned long long total = (unsigned long long)block_off[blockIdx.x] + block_cnt[blockIdx.x];
        cnt->n_lines = total;
        cnt->n_rows = 0;
        if (total > cap_lines) { cnt->err = ING_TOO_MANY_LINES; cnt->n_lines = cap_lines; }
    }
}

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }   // std::isspace, "C" locale

// line i = [previous newline + 1, newline i), trimmed like boost::trim (tped_genotype_file.cpp:116)
__global__ void line_span_kernel(const char *__restrict__ text, const uint32_t *__restrict__ nl, uint32_t cap_lines,
                                 const IngestCounters *__restrict__ cnt, uint2 *__restrict__ span, uint32_t *__restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap_lines) return;
    if (i >= cnt->n_lines) { flag[i] = 0; return; }
    uint32_t b = i ? nl[i - 1] + 1 : 0, e = nl[i];
    while (b < e && is_space(text[b])) ++b;
    while (e > b && is_space(text[e - 1])) --e;
    span[i] = make_uint2(b, e);
    flag[i] = e > b;
}

__device__ __forceinline__ int allele_index(char c) {   // tped_genotype_file.cpp:150-166 then the "ACGT" alphabet (table5.cpp:94-100)
    return (c == 'A' || c == '1') ? 0 : (c == 'C' || c == '2') ? 1 : (c == 'G' || c == '3') ? 2 : (c == 'T' || c == '4') ? 3 : 4;
}

// Warp-wide row assembler: lanes hand in one genotype encoding per step (enc < 0 = unknown), in column order.
struct RowAssembler {
    Labeler lab;
    uint32_t m1, m2, word;   // this lane's parked words of the current 32-word group; words finished so far
    uint32_t *p1, *p2;
    int bad_lane;            // first lane whose genotype the reference's state machine rejects, -1 = none

    __device__ void begin(uint32_t *row, uint32_t Wr) { lab.reset(); m1 = m2 = word = 0; p1 = row; p2 = row + Wr; bad_lane = -1; }

    __device__ void push32(int enc, uint32_t lane) {
        const bool known = enc >= 0;
        int c = known ? (int)((lab.codes >> (4 * enc)) & 0xF) : 0;
        uint32_t need = __ballot_sync(0xffffffffu, known && c == 0);
        while (need) {   // a new spelling: label it in column order (at most three per row)
            const int src = __ffs(need) - 1;
            const int e0 = __shfl_sync(0xffffffffu, enc, src);
            const int c0 = lab.code(e0);
            if (c0 < 0) { if (bad_lane < 0) bad_lane = src; lab.codes |= 0xFull << (4 * e0); }   // poison, keep going
            if (known && enc == e0) c = c0 < 0 ? 0 : c0;
            need = __ballot_sync(0xffffffffu, known && c == 0 && ((lab.codes >> (4 * enc)) & 0xF) == 0);
        }
        if (c == 0xF) c = 0;
        const uint32_t w1 = __ballot_sync(0xffffffffu, c & 1), w2 = __ballot_sync(0xffffffffu, c & 2);
        if (lane == (word & 31)) { m1 = w1; m2 = w2; }
        ++word;
        if ((word & 31) == 0) flush(lane, 32);
    }
    __device__ void flush(uint32_t lane, uint32_t n) {   // the last n parked words, 128-byte stores
        const uint32_t base = (word - 1) & ~31u;
        if (lane < n) { p1[base + lane] = m1; p2[base + lane] = m2; }
        m1 = m2 = 0;
    }
    __device__ void end(uint32_t Wr, uint32_t lane, uint16_t *hdr) {
        if (word & 31) flush(lane, word & 31);
        for (uint32_t w = word + lane; w < Wr; w += 32) { p1[w] = 0; p2[w] = 0; }
        if (lane == 0) *hdr = lab.head;
    }
};

__global__ void __launch_bounds__(256) tped_pack_kernel(const char *__restrict__ text, const uint2 *__restrict__ span,
                                                        const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
                                                        IngestCounters *cnt, uint32_t N, uint32_t Wr, uint64_t M,
                                                        uint16_t *__restrict__ hdr, uint32_t *__restrict__ raw) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_lines = cnt->n_lines, cursor = cnt->cursor;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t line = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < n_lines; line += n_warps) {
        if (line == n_lines - 1 && lane == 0) cnt->n_rows = (unsigned long long)rank[line] + flag[line];
        if (!flag[line]) continue;
        const uint64_t row = cursor + rank[line];
        if (row >= M) { if (lane == 0) atomicCAS(&cnt->err, ING_OK, ING_TOO_MANY_ROWS); continue; }
        const uint2 sp = span[line];
        // the four marker fields: chromosome, id, genetic distance, position (every delimiter counts, :118-130)
        uint32_t g0 = sp.y;
        int need = 4;
        for (uint32_t p = sp.x; p < sp.y && need > 0; p += 32) {
            const char c = p + lane < sp.y ? text[p + lane] : 'x';
            uint32_t m = __ballot_sync(0xffffffffu, c == ' ' || c == '\t');
            const int n = __popc(m);
            if (n >= need) {
                for (int k = 1; k < need; ++k) m &= m - 1;
                g0 = p + (__ffs(m) - 1) + 1;
                need = 0;
            } else need -= n;
        }
        if (need > 0) {
            if (lane == 0 && atomicCAS(&cnt->err, ING_OK, ING_SHORT_HEADER) == ING_OK) { cnt->err_row = row; cnt->err_col = 0; }
            continue;
        }
        const uint32_t ncols = min(N, (sp.y - g0 + 1) >> 2);   // "X Y" every 4 bytes (:138)
        RowAssembler ra;
        ra.begin(raw + row * 2ull * Wr, Wr);
        const char *g = text + g0;
        for (uint32_t c0 = 0; c0 < ncols; c0 += 128) {          // four steps' loads issued together
            int enc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t col = c0 + 32 * u + lane;
                enc[u] = -1;
                if (col < ncols) {
                    const int a = allele_index(g[4ull * col]), b = allele_index(g[4ull * col + 2]);
                    if (a < 4 && b < 4) enc[u] = 4 * a + b;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + 32 * u < ncols) {
                    ra.push32(enc[u], lane);
                    if (ra.bad_lane >= 0) {
                        if (lane == 0 && atomicCAS(&cnt->err, ING_OK, ING_THIRD_SPELLING) == ING_OK) { cnt->err_row = row; cnt->err_col = c0 + 32 * u + ra.bad_lane; }
                        ra.bad_lane = -2;
                    }
                }
        }
        ra.end(Wr, lane, hdr + row);
    }
}

__global__ void advance_cursor_kernel(IngestCounters *cnt) { cnt->cursor += cnt->n_rows; }

// PLINK .bed, SNP-major: ceil(N/4) bytes per SNP, genotype of sample c = bits 2(c&3).. of byte c>>2:
// 0 = homozygous A1, 1 = missing, 2 = heterozygous, 3 = homozygous A2. alleles[2r], alleles[2r+1] = index of A1, A2 in
// "ACGT" (from the .bim); the row gets the labels the text loader would give the same genotypes written "A1A1",
// "A1A2", "A2A2" in sample order.
__global__ void __launch_bounds__(256) bed_pack_kernel(const uint8_t *__restrict__ bed, uint32_t bytes_per_snp, const uint8_t *__restrict__ alleles,
                                                       uint64_t n_rows, uint64_t first_row, uint32_t N, uint32_t Wr,
                                                       uint16_t *__restrict__ hdr, uint32_t *__restrict__ raw) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += n_warps) {
        const int a1 = alleles ? alleles[2 * r] & 3 : 0, a2 = alleles ? alleles[2 * r + 1] & 3 : 1;
        const uint8_t *src = bed + r * (uint64_t)bytes_per_snp;
        RowAssembler ra;
        ra.begin(raw + (first_row + r) * 2ull * Wr, Wr);
        for (uint32_t c0 = 0; c0 < N; c0 += 32) {
            // 8 bytes hold this step's 32 genotypes; lanes 4q..4q+3 share byte q
            const uint32_t col = c0 + lane;
            int enc = -1;
            if (col < N) {
                const uint32_t g = (src[col >> 2] >> (2 * (col & 3))) & 3u;
                enc = g == 0 ? 5 * a1 : g == 2 ? 4 * a1 + a2 : g == 3 ? 5 * a2 : -1;
            }
            ra.push32(enc, lane);
        }
        ra.end(Wr, lane, hdr + first_row + r);
    }
}

static int ingest_get(gwasdev_store *s, Ingest **out) {
    if (!s->ingest) {
        Ingest *g = new Ingest();
        cudaError_t e = cudaMalloc(&g->d_cnt, sizeof(IngestCounters));
        if (e == cudaSuccess) e = cudaMallocHost((void **)&g->h_cnt, sizeof(IngestCounters));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done[0], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done[1], cudaEventDisableTiming);
        s->ingest = g;
        GW_CUDA(e);
    }
    *out = (Ingest *)s->ingest;
    return GWASDEV_OK;
}

template <class T> static cudaError_t grow(T *&p, size_t &cap, size_t n) {   // elements
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

// device text buffer b holds `len` bytes ending in '\n': index, rank and pack its lines at the device row cursor
static int ingest_text_chunk(gwasdev_store *s, Ingest *g, int b, size_t len) {
    const size_t padded = (len + NL_BLOCK - 1) / NL_BLOCK * NL_BLOCK;
    const uint32_t n_blocks = (uint32_t)(padded / NL_BLOCK);
    const uint32_t cap_lines = (uint32_t)std::min<uint64_t>(len, 2 * s->M + 64);
    if (n_blocks > g->cap_blocks) {
        size_t c1 = g->cap_blocks, c2 = g->cap_blocks;
        GW_CUDA(grow(g->d_block_cnt, c1, n_blocks));
        GW_CUDA(grow(g->d_block_off, c2, n_blocks));
        g->cap_blocks = n_blocks;
    }
    if (cap_lines > g->cap_lines) {
        size_t c1 = g->cap_lines, c2 = g->cap_lines, c3 = g->cap_lines, c4 = g->cap_lines;
        GW_CUDA(grow(g->d_nl, c1, cap_lines));
        GW_CUDA(grow(g->d_flag, c2, cap_lines));
        GW_CUDA(grow(g->d_rank, c3, cap_lines));
        GW_CUDA(grow(g->d_span, c4, cap_lines));
        g->cap_lines = cap_lines;
    }
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp1, g->d_block_cnt, g->d_block_off, (int)n_blocks, s->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, g->d_flag, g->d_rank, (int)cap_lines, s->stream);
    if (std::max(tmp1, tmp2) > g->cap_tmp) {
        if (g->d_tmp) cudaFree(g->d_tmp);
        g->d_tmp = nullptr; g->cap_tmp = 0;
        GW_CUDA(cudaMalloc(&g->d_tmp, std::max(tmp1, tmp2)));
        g->cap_tmp = std::max(tmp1, tmp2);
    }
    char *text = g->d_text[b];
    if (padded > len) GW_CUDA(cudaMemsetAsync(text + len, 0, padded - len, s->stream));
    newline_count_kernel<<<n_blocks, NL_THREADS, 0, s->stream>>>(reinterpret_cast<const uint4 *>(text), g->d_block_cnt);
    GW_LAUNCHED();
    size_t t = g->cap_tmp;
    GW_CUDA(cub::DeviceScan::ExclusiveSum(g->d_tmp, t, g->d_block_cnt, g->d_block_off, (int)n_blocks, s->stream));
    ++g_launches;
    newline_write_kernel<<<n_blocks, NL_THREADS, 0, s->stream>>>(reinterpret_cast<const uint4 *>(text), g->d_block_cnt, g->d_block_off,
                                                                 g->d_nl, cap_lines, g->d_cnt);
    GW_LAUNCHED();
    line_span_kernel<<<(cap_lines + 255) / 256, 256, 0, s->stream>>>(text, g->d_nl, cap_lines, g->d_cnt, g->d_span, g->d_flag);
    GW_LAUNCHED();
    t = g->cap_tmp;
    GW_CUDA(cub::DeviceScan::ExclusiveSum(g->d_tmp, t, g->d_flag, g->d_rank, (int)cap_lines, s->stream));
    ++g_launches;
    int sms = 0;
    GW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    tped_pack_kernel<<<sms * 8, 256, 0, s->stream>>>(text, g->d_span, g->d_flag, g->d_rank, g->d_cnt, s->N, s->Wr, s->M, s->d_hdr, s->d_raw);
    GW_LAUNCHED();
    advance_cursor_kernel<<<1, 1, 0, s->stream>>>(g->d_cnt);
    GW_LAUNCHED();
    return GWASDEV_OK;
}

static int ingest_finish(gwasdev_store *s, Ingest *g, const char *who, uint64_t first_row, uint64_t *rows_done) {
    GW_CUDA(cudaMemcpyAsync(g->h_cnt, g->d_cnt, sizeof(IngestCounters), cudaMemcpyDeviceToHost, s->stream));
    GW_CUDA(cudaStreamSynchronize(s->stream));
    const IngestCounters &c = *g->h_cnt;
    if (rows_done) *rows_done = c.cursor - first_row;
    switch (c.err) {
    case ING_OK: return GWASDEV_OK;
    case ING_THIRD_SPELLING:
        set_error("%s: row %llu, column %u introduces a third genotype spelling of one kind; the reference aborts here "
                  "(compressed_genotype_table5.cpp:325)", who, c.err_row, c.err_col);
        return GWASDEV_EINVAL;
    case ING_TOO_MANY_ROWS: set_error("%s: more genotype lines than the table has rows (%llu)", who, (unsigned long long)s->M); return GWASDEV_EINVAL;
    case ING_SHORT_HEADER: set_error("%s: row %llu has fewer than four marker fields", who, c.err_row); return GWASDEV_EINVAL;
    default: set_error("%s: more than %llu lines in one chunk (mostly blank input?)", who, 2ull * s->M + 64); return GWASDEV_EINVAL;
    }
}

static int ingest_begin(gwasdev_store *s, Ingest *g, uint64_t first_row) {
    memset(g->h_cnt, 0, sizeof(IngestCounters));
    g->h_cnt->cursor = first_row;
    GW_CUDA(cudaMemcpyAsync(g->d_cnt, g->h_cnt, sizeof(IngestCounters), cudaMemcpyHostToDevice, s->stream));
    GW_CUDA(cudaStreamSynchronize(s->stream));   // h_cnt is reused for the read-back
    return GWASDEV_OK;
}

static int reserve_text(Ingest *g, size_t bytes, bool pinned) {
    const size_t padded = (bytes + NL_BLOCK - 1) / NL_BLOCK * NL_BLOCK;
    if (padded > g->cap_text) {
        for (int b = 0; b < 2; ++b) { if (g->d_text[b]) cudaFree(g->d_text[b]); g->d_text[b] = nullptr; }
        g->cap_text = 0;
        for (int b = 0; b < 2; ++b) GW_CUDA(cudaMalloc((void **)&g->d_text[b], padded));
        g->cap_text = padded;
    }
    if (pinned && bytes > g->cap_pin) {
        for (int b = 0; b < 2; ++b) { if (g->h_pin[b]) cudaFreeHost(g->h_pin[b]); g->h_pin[b] = nullptr; }
        g->cap_pin = 0;
        for (int b = 0; b < 2; ++b) GW_CUDA(cudaMallocHost((void **)&g->h_pin[b], bytes));
        g->cap_pin = bytes;
    }
    return GWASDEV_OK;
}

static void invalidate(gwasdev_store *s) { gwasdev_internal_rows_changed(s); }

static size_t chunk_bytes(const gwasdev_store *s) {
    if (s->opt[GWASDEV_OPT_INGEST_CHUNK] >= 64) return (size_t)s->opt[GWASDEV_OPT_INGEST_CHUNK];
    return 4ull << 20;   // pinning host memory costs ~1.5 ms per MB: small buffers, grown only for lines that do not fit
}

// Plain files are read with read() straight into the caller's (pinned) buffer; gzip files (magic 1f 8b) through zlib.
struct TextFile {
    int fd = -1;
    gzFile gz = nullptr;
    bool open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        unsigned char magic[2] = {0, 0};
        const ssize_t got = ::pread(fd, magic, 2, 0);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            gz = gzdopen(fd, "rb");
            if (!gz) { ::close(fd); fd = -1; return false; }
            gzbuffer(gz, 1 << 20);
        }
        return true;
    }
    bool compressed() const { return gz != nullptr; }
    long long size() const { struct stat st; return (!gz && fd >= 0 && fstat(fd, &st) == 0) ? (long long)st.st_size : -1; }
    long long read(char *buf, size_t n) {   // bytes read, 0 at end of file, < 0 on error
        if (gz) return gzread(gz, buf, (unsigned)std::min<size_t>(n, 1u << 30));
        for (;;) { const ssize_t r = ::read(fd, buf, std::min<size_t>(n, 1u << 30)); if (r < 0 && errno == EINTR) continue; return r; }
    }
    void close() { if (gz) gzclose(gz); else if (fd >= 0) ::close(fd); gz = nullptr; fd = -1; }
    ~TextFile() { close(); }
};

static bool is_blank(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// genotype columns of a TPED line, counted like the reference sizes its row buffer (tped_genotype_file.cpp:132-136)
static uint32_t tped_line_columns(const char *line, size_t len) {
    size_t b = 0, e = len;
    while (b < e && is_blank(line[b])) ++b;
    while (e > b && is_blank(line[e - 1])) --e;
    int delims = 0;
    while (b < e && delims < 4) { if (line[b] == ' ' || line[b] == '\t') ++delims; ++b; }
    return delims == 4 ? (uint32_t)((e - b + 1) >> 2) : 0;
}

}  // namespace gwasdev

using namespace gwasdev;

void gwasdev_internal_free_ingest(gwasdev_store *s) {
    Ingest *g = (Ingest *)s->ingest;
    if (!g) return;
    cudaFree(g->d_cnt); cudaFreeHost(g->h_cnt);
    for (int b = 0; b < 2; ++b) { cudaFree(g->d_text[b]); if (g->h_pin[b]) cudaFreeHost(g->h_pin[b]); if (g->done[b]) cudaEventDestroy(g->done[b]); }
    cudaFree(g->d_block_cnt); cudaFree(g->d_block_off); cudaFree(g->d_nl); cudaFree(g->d_flag); cudaFree(g->d_rank); cudaFree(g->d_span);
    cudaFree(g->d_tmp); cudaFree(g->d_alleles);
    delete g;
    s->ingest = nullptr;
}

extern "C" {

int gwasdev_put_tped_text(gwasdev_store *s, uint64_t first_row, const char *text, size_t len, uint64_t *rows_done, size_t *bytes_used) {
    GW_REQUIRE(s && text, "gwasdev_put_tped_text: NULL argument");
    GW_REQUIRE(first_row <= s->M, "gwasdev_put_tped_text: first row %llu outside the table of %llu", (unsigned long long)first_row, (unsigned long long)s->M);
    GW_REQUIRE(len < (1ull << 32) - NL_BLOCK, "gwasdev_put_tped_text: chunks are limited to 4 GiB; call again with the rest");
    if (rows_done) *rows_done = 0;
    // whole lines only: the caller keeps what follows the last newline for its next call
    size_t used = len;
    while (used > 0 && text[used - 1] != '\n') --used;
    if (bytes_used) *bytes_used = used;
    if (used == 0) return GWASDEV_OK;
    GW_CUDA(cudaSetDevice(s->device));
    Ingest *g = nullptr;
    int rc = ingest_get(s, &g);
    if (rc == GWASDEV_OK) rc = reserve_text(g, used, false);
    if (rc == GWASDEV_OK) rc = ingest_begin(s, g, first_row);
    if (rc != GWASDEV_OK) return rc;
    GW_CUDA(cudaMemcpyAsync(g->d_text[0], text, used, cudaMemcpyHostToDevice, s->stream));
    rc = ingest_text_chunk(s, g, 0, used);
    if (rc != GWASDEV_OK) return rc;
    invalidate(s);
    return ingest_finish(s, g, "gwasdev_put_tped_text", first_row, rows_done);
}

int gwasdev_tped_dims(const char *path, uint64_t *n_rows, uint32_t *n_samples) {
    GW_REQUIRE(path && n_rows && n_samples, "gwasdev_tped_dims: NULL argument");
    TextFile f;
    GW_REQUIRE(f.open(path), "gwasdev_tped_dims: cannot open %s", path);
    std::vector<char> buf(8u << 20);
    std::string first;
    bool first_done = false, line_has_text = false;
    uint64_t rows = 0;
    long long n;
    while ((n = f.read(buf.data(), buf.size())) > 0) {
        for (long long i = 0; i < n;) {
            const char *nlp = (const char *)memchr(buf.data() + i, '\n', (size_t)(n - i));
            const long long e = nlp ? (long long)(nlp - buf.data()) : n;
            if (!first_done) first.append(buf.data() + i, (size_t)(e - i));
            if (!line_has_text)
                for (long long q = i; q < e; ++q) if (!is_blank(buf[q])) { line_has_text = true; break; }
            if (nlp) {
                rows += line_has_text;
                if (!first_done) { if (line_has_text) first_done = true; else first.clear(); }
                line_has_text = false;
                i = e + 1;
            }
            else i = n;
        }
    }
    GW_REQUIRE(n == 0, "gwasdev_tped_dims: read error in %s", path);
    rows += line_has_text;   // last line without a newline
    *n_rows = rows;
    *n_samples = tped_line_columns(first.data(), first.size());
    return GWASDEV_OK;
}

// the rest of an open text file (after `head`, bytes already read from it) into rows first_row.. of the store
static int load_text_file(gwasdev_store *s, TextFile &f, const char *path, const char *head, size_t head_len, uint64_t first_row,
                          uint64_t *rows_done) {
    size_t CH = std::max(chunk_bytes(s), head_len + 1);
    Ingest *g = nullptr;
    int rc = ingest_get(s, &g);
    if (rc == GWASDEV_OK) rc = reserve_text(g, CH + 1, true);
    if (rc == GWASDEV_OK) rc = ingest_begin(s, g, first_row);
    if (rc != GWASDEV_OK) return rc;
    // two pinned buffers: the file is read into one while the device works on the other
    size_t carry = head_len;
    if (head_len) memcpy(g->h_pin[0], head, head_len);
    int b = 0;
    bool eof = false, used_buf[2] = {false, false};
    while (!eof || carry) {
        if (used_buf[b]) GW_CUDA(cudaEventSynchronize(g->done[b]));
        char *buf = g->h_pin[b];
        size_t have = carry;
        while (!eof && have < CH) {
            const long long n = f.read(buf + have, CH - have);
            GW_REQUIRE(n >= 0, "gwasdev_load_tped: read error in %s", path);
            if (n == 0) eof = true;
            have += (size_t)n;
        }
        if (eof && have > 0 && buf[have - 1] != '\n') buf[have++] = '\n';   // last line without a newline (room: CH + 1)
        size_t used = have;
        while (used > 0 && buf[used - 1] != '\n') --used;
        if (used == 0) {
            if (have == 0) break;
            // a line longer than the buffers: quadruple them (the device may still be reading the other one) and go on
            GW_REQUIRE(CH < (1ull << 30), "gwasdev_load_tped: a line of %s is longer than 1 GiB", path);
            GW_CUDA(cudaStreamSynchronize(s->stream));
            std::vector<char> keep(buf, buf + have);
            CH *= 4;
            if ((rc = reserve_text(g, CH + 1, true)) != GWASDEV_OK) return rc;
            memcpy(g->h_pin[b], keep.data(), have);
            used_buf[0] = used_buf[1] = false;
            carry = have;
            continue;
        }
        GW_CUDA(cudaMemcpyAsync(g->d_text[b], buf, used, cudaMemcpyHostToDevice, s->stream));
        if ((rc = ingest_text_chunk(s, g, b, used)) != GWASDEV_OK) return rc;
        GW_CUDA(cudaEventRecord(g->done[b], s->stream));
        used_buf[b] = true;
        carry = have - used;
        if (carry) {
            if (used_buf[b ^ 1]) GW_CUDA(cudaEventSynchronize(g->done[b ^ 1]));
            memcpy(g->h_pin[b ^ 1], buf + used, carry);
        }
        b ^= 1;
    }
    invalidate(s);
    return ingest_finish(s, g, "gwasdev_load_tped", first_row, rows_done);
}

// Plain files: the file is mapped and handed to the driver chunk by chunk as pageable memory (it stages such copies through
// its own pinned buffers at ~10 GB/s, the speed of read()); no pinned allocation of our own -- which costs ~1.5 ms per MB
// and dominated a cold load -- and no second copy. Chunks end on a newline; a last line without one gets it on the device.
static int load_mapped_file(gwasdev_store *s, int fd, size_t size, const char *path, uint64_t first_row, uint64_t *rows_done) {
    Ingest *g = nullptr;
    int rc = ingest_get(s, &g);
    if (rc == GWASDEV_OK) rc = ingest_begin(s, g, first_row);
    if (rc != GWASDEV_OK) return rc;
    if (size == 0) return ingest_finish(s, g, "gwasdev_load_tped", first_row, rows_done);
    const char *map = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE | (size <= (1ull << 30) ? MAP_POPULATE : 0), fd, 0);   // small files: no page fault per 4 KiB
    GW_REQUIRE(map != MAP_FAILED, "gwasdev_load_tped: cannot map %s", path);
    madvise((void *)map, size, MADV_SEQUENTIAL);
    const size_t CH = s->opt[GWASDEV_OPT_INGEST_CHUNK] ? chunk_bytes(s) : (32ull << 20);
    int b = 0;
    for (size_t off = 0; off < size && rc == GWASDEV_OK;) {
        size_t end = std::min(size, off + CH);
        if (end < size) {   // back to the last newline of the window; a line longer than the window runs to its own end
            const char *nl = (const char *)memrchr(map + off, '\n', end - off);
            if (!nl) nl = (const char *)memchr(map + end, '\n', size - end);
            end = nl ? (size_t)(nl - map) + 1 : size;
        }
        const size_t len = end - off;
        const bool add_newline = map[end - 1] != '\n';     // only the file's last line can lack it
        if (len + 1 >= (1ull << 32) - NL_BLOCK) { set_error("gwasdev_load_tped: a line of %s is longer than 4 GiB", path); rc = GWASDEV_EINVAL; break; }
        if ((rc = reserve_text(g, len + 1, false)) != GWASDEV_OK) break;   // grows (after a stream synchronisation inside cudaFree) for long lines
        cudaError_t e = cudaMemcpyAsync(g->d_text[b], map + off, len, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess && add_newline) e = cudaMemsetAsync(g->d_text[b] + len, '\n', 1, s->stream);
        if (e != cudaSuccess) { set_error("gwasdev_load_tped: %s", cudaGetErrorString(e)); rc = GWASDEV_ENODEVICE; break; }
        rc = ingest_text_chunk(s, g, b, len + (add_newline ? 1 : 0));
        b ^= 1;
        off = end;
    }
    if (rc != GWASDEV_OK) cudaStreamSynchronize(s->stream);
    munmap((void *)map, size);
    if (rc != GWASDEV_OK) return rc;
    invalidate(s);
    return ingest_finish(s, g, "gwasdev_load_tped", first_row, rows_done);
}

int gwasdev_load_tped(gwasdev_store *s, const char *path, uint64_t first_row, uint64_t *rows_done) {
    GW_REQUIRE(s && path, "gwasdev_load_tped: NULL argument");
    GW_REQUIRE(first_row <= s->M, "gwasdev_load_tped: first row outside the table");
    if (rows_done) *rows_done = 0;
    GW_CUDA(cudaSetDevice(s->device));
    TextFile f;
    GW_REQUIRE(f.open(path), "gwasdev_load_tped: cannot open %s", path);
    if (!f.compressed() && f.size() >= 0) return load_mapped_file(s, f.fd, (size_t)f.size(), path, first_row, rows_done);
    return load_text_file(s, f, path, nullptr, 0, first_row, rows_done);
}

int gwasdev_create_from_tped(const char *path, int device, gwasdev_store **out, uint64_t *n_rows, uint32_t *n_samples) {
    GW_REQUIRE(path && out, "gwasdev_create_from_tped: NULL argument");
    *out = nullptr;
    uint64_t rows_cap = 0;
    uint32_t cols = 0;
    TextFile f;
    GW_REQUIRE(f.open(path), "gwasdev_create_from_tped: cannot open %s", path);
    if (f.compressed()) {   // size unknown before inflating: the counting pass, then a second open
        f.close();
        int rc = gwasdev_tped_dims(path, &rows_cap, &cols);
        if (rc != GWASDEV_OK) return rc;
        GW_REQUIRE(f.open(path), "gwasdev_create_from_tped: cannot open %s", path);
    } else {
        // one pass: the first non-blank line gives the sample count, the file size an upper bound of the rows (a line is at
        // least 8 bytes of marker fields and 4 bytes per sample); the table is trimmed to the rows actually found
        const long long size = f.size();
        std::vector<char> buf(1u << 16);
        long long pos = 0;
        std::string line;
        while (cols == 0 && pos < size) {   // first line that parses (pread: the mapping is made by the loader)
            const ssize_t n = ::pread(f.fd, buf.data(), buf.size(), pos);
            GW_REQUIRE(n >= 0, "gwasdev_create_from_tped: read error in %s", path);
            if (n == 0) break;
            for (ssize_t i = 0; i < n && cols == 0; ++i) {
                if (buf[i] != '\n') { line.push_back(buf[i]); continue; }
                cols = tped_line_columns(line.data(), line.size());
                line.clear();
            }
            pos += n;
        }
        if (cols == 0) cols = tped_line_columns(line.data(), line.size());
        if (cols) rows_cap = (uint64_t)size / (4ull * cols + 7) + 1;
    }
    GW_REQUIRE(cols > 0 && rows_cap > 0, "gwasdev_create_from_tped: %s holds no genotype line with four marker fields", path);
    gwasdev_store *s = nullptr;
    int rc = gwasdev_create(rows_cap, cols, device, &s);
    if (rc != GWASDEV_OK) return rc;
    uint64_t rows = 0;
    rc = f.compressed() ? load_text_file(s, f, path, nullptr, 0, 0, &rows) : load_mapped_file(s, f.fd, (size_t)f.size(), path, 0, &rows);
    if (rc == GWASDEV_OK && rows == 0) { set_error("gwasdev_create_from_tped: %s holds no genotype rows", path); rc = GWASDEV_EINVAL; }
    if (rc != GWASDEV_OK) { gwasdev_destroy(s); return rc; }
    s->M = rows;                                   // rows beyond are allocated but not part of the table
    s->Mpad = (rows + TILE - 1) / TILE * TILE;
    *out = s;
    if (n_rows) *n_rows = rows;
    if (n_samples) *n_samples = cols;
    return GWASDEV_OK;
}

int gwasdev_put_bed(gwasdev_store *s, uint64_t first_row, uint64_t n_rows, const uint8_t *bed, const uint8_t *alleles) {
    GW_REQUIRE(s && bed, "gwasdev_put_bed: NULL argument");
    GW_REQUIRE(first_row + n_rows <= s->M, "gwasdev_put_bed: rows [%llu, %llu) outside the table of %llu", (unsigned long long)first_row,
               (unsigned long long)(first_row + n_rows), (unsigned long long)s->M);
    if (alleles)
        for (uint64_t r = 0; r < n_rows; ++r)
            GW_REQUIRE(alleles[2 * r] < 4 && alleles[2 * r + 1] < 4 && alleles[2 * r] != alleles[2 * r + 1],
                       "gwasdev_put_bed: row %llu needs two different alleles out of ACGT (indices 0-3)", (unsigned long long)(first_row + r));
    if (n_rows == 0) return GWASDEV_OK;
    GW_CUDA(cudaSetDevice(s->device));
    Ingest *g = nullptr;
    int rc = ingest_get(s, &g);
    if (rc != GWASDEV_OK) return rc;
    const uint32_t bps = (s->N + 3) / 4;
    const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / bps);
    GW_CUDA(reserve(s->sc_stage, std::min(chunk, n_rows) * bps));
    if (alleles) GW_CUDA(grow(g->d_alleles, g->cap_alleles, (size_t)(2 * std::min(chunk, n_rows))));
    int sms = 0;
    GW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    for (uint64_t r = 0; r < n_rows; r += chunk) {
        const uint64_t n = std::min(chunk, n_rows - r);
        GW_CUDA(cudaMemcpyAsync(s->sc_stage.p, bed + r * bps, n * bps, cudaMemcpyHostToDevice, s->stream));
        if (alleles) GW_CUDA(cudaMemcpyAsync(g->d_alleles, alleles + 2 * r, 2 * n, cudaMemcpyHostToDevice, s->stream));
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)sms * 8);
        bed_pack_kernel<<<grid, 256, 0, s->stream>>>((const uint8_t *)s->sc_stage.p, bps, alleles ? g->d_alleles : nullptr, n, first_row + r,
                                                     s->N, s->Wr, s->d_hdr, s->d_raw);
        GW_LAUNCHED();
        GW_CUDA(cudaStreamSynchronize(s->stream));
    }
    invalidate(s);
    return GWASDEV_OK;
}

int gwasdev_bed_dims(const char *bed_path, uint32_t n_samples, uint64_t *n_rows) {
    GW_REQUIRE(bed_path && n_rows && n_samples > 0, "gwasdev_bed_dims: bad argument");
    FILE *f = fopen(bed_path, "rb");
    GW_REQUIRE(f != nullptr, "gwasdev_bed_dims: cannot open %s", bed_path);
    unsigned char magic[3] = {0, 0, 0};
    const size_t got = fread(magic, 1, 3, f);
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fclose(f);
    GW_REQUIRE(got == 3 && magic[0] == 0x6c && magic[1] == 0x1b, "gwasdev_bed_dims: %s is not a PLINK .bed file", bed_path);
    GW_REQUIRE(magic[2] == 1, "gwasdev_bed_dims: %s is individual-major; only SNP-major .bed files are supported", bed_path);
    const uint64_t bps = (n_samples + 3) / 4;
    GW_REQUIRE((uint64_t)(size - 3) % bps == 0, "gwasdev_bed_dims: %s does not hold whole rows of %u samples", bed_path, n_samples);
    *n_rows = (uint64_t)(size - 3) / bps;
    return GWASDEV_OK;
}

int gwasdev_load_bed(gwasdev_store *s, const char *bed_path, const uint8_t *alleles, uint64_t first_row, uint64_t *rows_done) {
    GW_REQUIRE(s && bed_path, "gwasdev_load_bed: NULL argument");
    if (rows_done) *rows_done = 0;
    uint64_t rows = 0;
    int rc = gwasdev_bed_dims(bed_path, s->N, &rows);
    if (rc != GWASDEV_OK) return rc;
    GW_REQUIRE(first_row + rows <= s->M, "gwasdev_load_bed: %llu rows do not fit the table of %llu from row %llu", (unsigned long long)rows,
               (unsigned long long)s->M, (unsigned long long)first_row);
    // the file is mapped and its rows go to the device in 64 MB pieces as pageable memory (staged by the driver)
    const int fd = ::open(bed_path, O_RDONLY);
    GW_REQUIRE(fd >= 0, "gwasdev_load_bed: cannot open %s", bed_path);
    const uint64_t bps = (s->N + 3) / 4, size = 3 + rows * bps;
    const uint8_t *map = rows ? (const uint8_t *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE | (size <= (1ull << 30) ? MAP_POPULATE : 0), fd, 0) : nullptr;
    ::close(fd);
    GW_REQUIRE(rows == 0 || map != MAP_FAILED, "gwasdev_load_bed: cannot map %s", bed_path);
    if (rows) {
        rc = gwasdev_put_bed(s, first_row, rows, map + 3, alleles);
        munmap((void *)map, size);
        if (rc != GWASDEV_OK) return rc;
    }
    if (rows_done) *rows_done = rows;
    return GWASDEV_OK;
}

}  // extern "C"
