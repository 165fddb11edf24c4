// Pieces shared by the two pair-screen engines (pairwise.cu: AND+POPC tiles; pairwise_mma.cu: tcgen05 tiles).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gwasdev {

struct Candidate { uint32_t i, j; float stat; uint32_t pad; };

// Where the screen kernels put the pairs that pass their fp32 test. Threshold mode (hist == nullptr): every pair above
// `thr` (= threshold - margin) is appended; the host re-runs with a larger buffer if it overflowed. Top-k mode: the
// kernels also keep a histogram of the appended statistics, and whenever the buffer passes a fill mark one thread
// raises a device-wide threshold to the lower edge of the bin above which k candidates have already been seen (minus
// `slack`, twice the fp32 error budget: the final choice is made on the fp64 re-scores) -- a pair below it can no
// longer be among the k best, so the buffer stops growing however low the caller's threshold is.
constexpr int CAND_HIST_BINS = 2048;
struct CandSink {
    Candidate *cand;
    unsigned long long *n_cand;       // pairs appended so far (may exceed cap)
    unsigned long long cap;
    float thr;                        // static floor
    uint32_t *hist;                   // [CAND_HIST_BINS] of width hist_w from thr up, last bin open-ended; nullptr: threshold mode
    float *thr_dyn;                   // current device-wide threshold
    int *lost_max;                    // largest statistic (float bits, >= 0) among the pairs that found the buffer full
    float hist_w, slack;
    unsigned long long k_keep;
    unsigned long long raise_from, raise_mask;   // raise when slot + 1 >= raise_from and ((slot + 1) & raise_mask) == 0
};

// threshold the pairs of the next tile are tested against
__device__ __forceinline__ float sink_threshold(const CandSink &s) {
    if (!s.hist) return s.thr;
    float t;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(t) : "l"(s.thr_dyn) : "memory");
    return fmaxf(t, s.thr);
}
static __device__ __noinline__ void sink_raise(const uint32_t *hist, float *thr_dyn, float thr, float hist_w, float slack, unsigned long long k_keep) {
    unsigned long long acc = 0;
    for (int b = CAND_HIST_BINS - 1; b >= 0; --b) {
        acc += *reinterpret_cast<const volatile uint32_t *>(hist + b);
        if (acc >= k_keep) {
            const float t = thr + (float)b * hist_w - slack;
            if (t > thr && t > 0.f) atomicMax(reinterpret_cast<int *>(thr_dyn), __float_as_int(t));   // positive floats order like ints
            return;
        }
    }
}
// append pair (i, j); the caller has checked stat > sink_threshold
__device__ __forceinline__ void sink_push(const CandSink &s, uint32_t i, uint32_t j, float stat) {
    if (s.hist) {
        const int b = min(max((int)((stat - s.thr) / s.hist_w), 0), CAND_HIST_BINS - 1);
        atomicAdd(s.hist + b, 1u);
    }
    const unsigned long long slot = atomicAdd(s.n_cand, 1ull);
    if (slot < s.cap) { Candidate cd; cd.i = i; cd.j = j; cd.stat = stat; cd.pad = 0; s.cand[slot] = cd; }
    else if (s.hist) atomicMax(s.lost_max, __float_as_int(fmaxf(stat, 0.f)));
    if (s.hist && slot + 1 >= s.raise_from && ((slot + 1) & s.raise_mask) == 0) sink_raise(s.hist, s.thr_dyn, s.thr, s.hist_w, s.slack, s.k_keep);
}

// ---- PTX helpers (mbarrier + TMA) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}


typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
int get_encode_tiled(encode_tiled_fn *out);

__device__ __forceinline__ float u2f(uint32_t n) { return __uint_as_float(0x4B000000u | n) - 8388608.0f; }   // n < 2^23

// fp32 KSA screen value of a 3x3x2 table with the per-SNP records of both SNPs (see pairwise.cu for the algebra)
__device__ __forceinline__ float g_nlogn(float n) { return n * __logf(fmaxf(n, 1.0f)); }

__device__ __forceinline__ float ksa_screen_f32(const uint32_t (&n)[2][3][3], const PairSide &A, const PairSide &B,
                                                float N, float lnN) {
    float S = 0.f, tau = 0.f, total = 0.f;
    float row[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, col[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float c0 = u2f(n[0][a][b]), c1 = u2f(n[1][a][b]), cab = c0 + c1;
            const float W = fmaf(B.w[0][b], A.pca[0][a], B.w[1][b] * A.pca[1][a]);
            tau = fmaf(cab, W, tau);
            S += g_nlogn(c0) + g_nlogn(c1) - g_nlogn(cab);
            row[0][a] += c0; row[1][a] += c1; col[0][b] += c0; col[1][b] += c1;
            total += cab;
        }
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            S = fmaf(-row[k][g], A.lpca[k][g], S);
            S = fmaf(-col[k][g], B.lw[k][g], S);
        }
    S = fmaf(-total, lnN, S);
    return 2.0f * fmaf(N, logf(tau), S);
}


}  // namespace gwasdev
