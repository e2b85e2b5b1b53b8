// dto_device.cuh -- device-side building blocks of the DTO hot path (sm_100a).
//
// What the reference computes per (t1, t2) cell (src/dto/process_threshold_pairs.rs:84-128):
//   k = |set1(t1) ∩ set2(t2)|                      (src/stat_operations/intersect_genes.rs:38-56)
//   p = hypergeometric_pvalue(N, |set1|, |set2|, k) (src/stat_operations/hypergeometric_pvalue.rs:33-50,
//                                                    statrs 0.17.1 Hypergeometric::sf)
// The functions here restate that arithmetic for the device in the SAME operation order as statrs, reading
// ln_factorial from a table the host builds with statrs' own recipe (dto_host_math.hpp).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/dto_b200.h"

namespace dto {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint16_t kNoSlot = 0xFFFFu;  // "no partner / partner beyond the last threshold"
constexpr int kMaxLevels = 32;         // screen levels 1..kMaxLevels: tau_1 = 0.95, tau_l = 0.5 * 2^-((l-2)/2); level 0 = "valid cell"

// log-domain safety margin: covers rounding of log pmf (~1e-10 abs) and of the statrs sum (~1e-12 rel)
constexpr double kEps = 1e-6;
// exp(x) underflows to exactly 0.0 in binary64 for x < -745.1332191019412.  A cell whose LARGEST tail term
// lies below kZeroLo has p == 0.0 in the reference; between kZeroLo and kZeroHi it is evaluated exactly.
// the ratio-recurrence tail reproduces log p of the reference to ~1e-10 (both sum the same ln-factorial table, whose
// 7-term log pmf carries ~1e-10 absolute rounding); cells within kRefineEps of the minimum go to the exact stage
constexpr double kRefineEps = 1e-8;      // floor of Problem::refine_eps (which grows with ulp(lf[N]) for huge populations)
constexpr uint32_t kRefinedBit = 0x80000000u;   // Cand::k flag: `v` is a log-p bound good to refine_eps
constexpr uint32_t kEvalBit = 0x40000000u;      // Cand::k flag: `v` is the exact (statrs-order) p-value itself
constexpr uint32_t kOverlapMask = 0x0000FFFFu;  // Cand::k: the overlap count
constexpr double kZeroLo = -745.14;
constexpr double kZeroHi = -745.12;

// Problem description resident in HBM; pointers are device pointers. Passed to kernels by value.
struct Problem {
    int T1, T2;
    int CH;          // columns per lane in the scan kernel (template value actually used)
    int CHP;         // hist_words(CH): 32-bit words per lane of the per-warp row histogram (two 16-bit column counts per
                     // word, padded to whole 16-byte vectors with an odd vector count so 128-bit reads are conflict-free)
    int T2pad;       // u16 entries per row of the critical-overlap tables: kcrit_row_u16(CH)
    int levels;      // number of screen levels beyond level 0
    uint32_t never;  // kcrit value meaning "no overlap passes": 0x7FFF when every set size is <= 32766 (packed 15-bit
                     // screen), else 0xFFFF
    uint32_t n1, n2;
    uint32_t n1_eff;     // #list-1 positions whose rank is <= the last threshold of list 1
    uint32_t pb_stride;  // u16 elements per permutation row of the partner-bin array
    uint32_t n_common;   // genes present in both lists
    uint64_t N;          // population
    const uint32_t *c1;  // [T1] set1_len per threshold  (K_i) == row end positions
    const uint32_t *c2;  // [T2] set2_len per threshold  (n_j)
    const uint32_t *thr1, *thr2;
    const double *lf;    // [N+1] ln_factorial
    const double *rowA;  // [T1] lf[K] + lf[N-K]
    const double *colB;  // [T2] lf[n] + lf[N-n] - lf[N]
    const uint16_t *kcrit;     // [(levels+1)][T1][T2pad], column j stored at kcrit_col(j, CH)
    const uint2 *cellmeta;     // [T1*T2] {offset into lptab, kbase | count << 16}: log p tabulated for k in [kbase, kbase+count)
    const double *lptab;       // log p (ratio-recurrence tail, ~1e-10) for every (cell, k) between tau_1 and tau_levels
    const uint16_t *dslot2;    // [n2, padded to a multiple of 8] row-histogram slot of list-2 position (or kNoSlot)
    const uint16_t *bin1;      // [n1] threshold bin of list-1 position (or kNoSlot)
    const uint16_t *bin2;      // [n2]
    const int32_t *slot2_of_1; // [n1]
    double level_log[kMaxLevels + 1];  // log tau_l ; [0] = +inf
    // Error budget of everything derived from log pmf = a 7-term sum of ln-factorials: max(1e-8, 40 ulp(lf[N])).  At
    // N = 20 000 (lf[N] ~ 1.8e5, ulp 2.9e-11) the floor applies; the accepted population limit 2^27 has ulp 4.8e-7.
    double refine_eps;
    double tab_slack;  // relative slack of the critical-overlap tables against the same rounding: max(1e-6, 4 refine_eps)
};

// One cell of a tie set the device could not settle (more than one distinct (K, n, k) inside the ambiguity window of the
// minimum): re-evaluated on the host with the host libm, dto_engine.cu.
struct TieEntry {
    uint32_t task;
    uint32_t ij;  // row << 16 | column
    uint32_t k;
};

// What one scan launch reports besides the records (16 B, read back once per launch instead of a status word per task).
struct ScanSummary {
    uint32_t n_done;         // tasks that reached the end of finish_task
    uint32_t n_full;         // tasks handed to the dense path (ids in full_list)
    uint32_t n_tie_entries;  // TieEntry slots requested (> tie_cap: the overflowing tasks went to full_list instead)
    uint32_t pad;
};

struct ScanOut {
    ScanSummary *summary;
    uint32_t *full_list;  // [n_tasks]
    TieEntry *ties;       // [tie_cap]
    uint32_t tie_cap;
};

// Row histogram of the scan kernel: lane = j / CH owns columns m = j % CH; columns (2q, 2q+1) of a lane share the
// 32-bit word q of that lane (low / high half).  hist_words = words per lane: CH/2 rounded up to whole uint4 vectors,
// and to an ODD number of vectors so that the 8 lanes of one 128-bit shared-memory wavefront cover all 32 banks.
__host__ __device__ constexpr int hist_words(int CH) {
    int v = (CH / 2 + 3) / 4;
    if (v % 2 == 0) ++v;
    return 4 * v;
}
// Histogram slot of column j as carried in the partner-slot rows: (byte offset of the word) | half, i.e.
// (word index << 2) | (j & 1); at most 32 * 36 words, so it fits 16 bits and never equals kNoSlot.
__host__ __device__ inline uint32_t hist_slot(int j, int CH) {
    const int lane = j / CH, m = j % CH;
    return ((uint32_t)(lane * hist_words(CH) + (m >> 1)) << 2) | (uint32_t)(m & 1);
}
__host__ __device__ inline uint32_t hist_slot_column(uint32_t slot, int CH) {
    const uint32_t word = slot >> 2, W = (uint32_t)hist_words(CH);
    return (word / W) * (uint32_t)CH + 2u * (word % W) + (slot & 1u);
}

// Position of column j inside a kcrit row (u16 units): lane = j / CH owns columns m = j % CH; its pair q = m / 2 is one
// 32-bit word; the lane's pairs 4v .. 4v+3 form one 16-byte vector, and vector v of all 32 lanes is contiguous, so a warp
// reads a row with ceil(CH/8) perfectly coalesced 128-bit loads (512 contiguous bytes each).  Words of pairs >= CH/2 are
// padding and hold "never".
__host__ __device__ constexpr int kcrit_vecs(int CH) { return (CH / 2 + 3) / 4; }
__host__ __device__ constexpr int kcrit_row_u16(int CH) { return kcrit_vecs(CH) * 32 * 4 * 2; }
__host__ __device__ inline size_t kcrit_col(int j, int CH) {
    const int lane = j / CH, m = j % CH, q = m >> 1;
    return (((size_t)(q >> 2) * 32 + (size_t)lane) * 4 + (size_t)(q & 3)) * 2 + (size_t)(m & 1);
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: the permutation for (seed, perm id) is a pure
// function of its id, so results do not depend on batching or on the GPU a shard lands on.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ inline void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// ---------------------------------------------------------------------------------------------------
// statrs-order upper tail.  Preconditions (checked by callers): K <= N, n <= N.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ln_binomial_tab(const double *__restrict__ lf, uint64_t a, uint64_t b) {
    // statrs factorial::ln_binomial: -inf if b > a, else lf(a) - lf(b) - lf(a-b), left to right
    if (b > a) return -CUDART_INF;
    return __dsub_rn(__dsub_rn(lf[a], lf[b]), lf[a - b]);
}

// exp() as the reference's libm rounds it.  For normal results CUDA's exp (<= 1 ulp) is used as is.  Results in
// the SUBNORMAL range carry only a few bits, so "1 ulp" there is a large relative error and decides whether a
// term (hence a whole p-value) is exactly 0.0 -- which in turn decides the tie-break on the underflow plateau
// (optimize_main.rs:77-80 compares p with ==).  glibc rounds exp(x) ONCE onto the 2^-1074 grid (e_exp.c
// specialcase); do the same: q = exp(x + 1074 ln 2) computed with a two-part ln 2, rounded to nearest-even.
__device__ __forceinline__ double exp_like_libm(double x) {
    if (x >= -708.0) return exp(x);
    if (x < -746.0) return 0.0;
    const double ln2_hi = 6.93147180369123816490e-01;  // 32 significant bits: 1074 * ln2_hi is exact
    const double ln2_lo = 1.90821492927058770002e-10;
    const double t = __dadd_rn(__dadd_rn(x, 1074.0 * ln2_hi), 1074.0 * ln2_lo);
    const double q = exp(t);  // in [~0.27, 2^52.6)
    return __longlong_as_double(__double2ll_rn(q));  // the subnormal r * 2^-1074 has bit pattern r
}

// hypergeometric_pvalue(N, K, n, k): k == 0 -> 1; sf(k-1): x < min -> 1, x >= max -> 0, else ascending
// sum_{i=k}^{min(K,n)} exp(lnC(K,i) + lnC(N-K,n-i) - lnC(N,n)).  Early exit is bit-preserving: past the mode
// terms fall monotonically and a term below 2^-55 of the accumulator cannot change it (nor can later ones).
__device__ inline double hypergeom_pvalue_exact(const double *__restrict__ lf, uint64_t N, uint64_t K, uint64_t n,
                                                uint64_t k) {
    if (K > N || n > N) return CUDART_NAN;
    if (k == 0) return 1.0;
    const uint64_t x = k - 1;
    const uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    const uint64_t mx = K < n ? K : n;
    if (x < mn) return 1.0;
    if (x >= mx) return 0.0;
    const double ln_denom = ln_binomial_tab(lf, N, n);
    const uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    const double lfK = lf[K], lfNK = lf[N - K];
    const uint64_t NK = N - K;
    double acc = 0.0;
    for (uint64_t i = x + 1; i <= mx; ++i) {
        const double a = __dsub_rn(__dsub_rn(lfK, lf[i]), lf[K - i]);
        const uint64_t ni = n - i;
        const double b = (ni > NK) ? -CUDART_INF : __dsub_rn(__dsub_rn(lfNK, lf[ni]), lf[NK - ni]);
        const double term = exp_like_libm(__dsub_rn(__dadd_rn(a, b), ln_denom));
        acc = __dadd_rn(acc, term);
        if (i > mode + 1 && term <= acc * 0x1p-55) break;
    }
    return acc;
}

// natural log of the same tail by running log-sum-exp (finite where the p-value underflows)
__device__ inline double hypergeom_log_pvalue(const double *__restrict__ lf, uint64_t N, uint64_t K, uint64_t n,
                                              uint64_t k) {
    if (K > N || n > N) return CUDART_NAN;
    if (k == 0) return 0.0;
    const uint64_t x = k - 1;
    const uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    const uint64_t mx = K < n ? K : n;
    if (x < mn) return 0.0;
    if (x >= mx) return -CUDART_INF;
    const double ln_denom = ln_binomial_tab(lf, N, n);
    const uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    double M = -CUDART_INF, S = 0.0;
    for (uint64_t i = x + 1; i <= mx; ++i) {
        const double a = ln_binomial_tab(lf, K, i) + ln_binomial_tab(lf, N - K, n - i) - ln_denom;
        if (a > M) {
            S = S * exp(M - a) + 1.0;
            M = a;
        } else {
            S += exp(a - M);
        }
        if (i > mode + 1 && a < M - 46.0) break;
    }
    return M + log(S);
}

// log pmf(k) for a valid cell (k > max(0, K+n-N)), using the per-row / per-column constants
__device__ __forceinline__ double log_pmf(const Problem &P, double rowA, double colB, uint32_t K, uint32_t n,
                                          uint32_t k) {
    const double *__restrict__ lf = P.lf;
    return rowA + colB - lf[k] - lf[K - k] - lf[n - k] - lf[P.N - K - n + k];
}

// lexicographic "better" of the reference's reduction (src/dto/optimize_main.rs:73-116):
// smaller p; then larger intersection; then smaller (rank1, rank2) == smaller (i, j)
struct Best {
    double p;
    uint32_t k;
    uint32_t ij;  // i << 16 | j ; 0xFFFFFFFF = none
};

__device__ __forceinline__ bool better(const Best &a, const Best &b) {
    if (a.ij == 0xFFFFFFFFu) return false;
    if (b.ij == 0xFFFFFFFFu) return true;
    if (a.p != b.p) return a.p < b.p;
    if (a.k != b.k) return a.k > b.k;
    return a.ij < b.ij;
}

__device__ __forceinline__ Best warp_best(Best v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.p = __shfl_xor_sync(kFull, v.p, off);
        o.k = __shfl_xor_sync(kFull, v.k, off);
        o.ij = __shfl_xor_sync(kFull, v.ij, off);
        if (better(o, v)) v = o;
    }
    return v;
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, off));
    return v;
}

}  // namespace dto
