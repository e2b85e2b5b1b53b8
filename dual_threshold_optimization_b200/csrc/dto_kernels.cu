// dto_kernels.cu -- hand-written sm_100a kernels of the DTO hot path.
//
//   K0  sigma_sort_kernel      one CTA per permutation: exactly-uniform random pairing of list-1 and list-2
//                              positions (Philox keys, shared-memory bucket sort), emitted as the row-histogram
//                              slot of every list-1 position's partner            (collections/permuted.rs:56-60,90-101)
//   K0' compose_pairing_kernel same output from HOST-supplied perm1/perm2 (parity mode)
//   K1  scan_kernel            one WARP per permutation: row-streamed histogram (shared-memory atomics) ->
//                              2-D inclusive prefix sum in registers/shuffles -> critical-overlap screen ->
//                              statrs-order FP64 tail for the few surviving cells -> warp-shuffle argmin with the
//                              reference tie-break        (process_threshold_pairs.rs:84-128, optimize_main.rs:73-116)
//   K2  build_kcrit_kernel     per-problem screen tables: smallest k with p(K_i, n_j, k) <= 4^-l
//   K3  full_* kernels         dense T1 x T2 grid (debug=true analogue, optimize_main.rs:68-70) + exact argmin;
//                              also the last-resort path for degenerate tasks (min p >= 1)
#include "dto_kernels.cuh"
#include "dto_host_math.hpp"  // kTieRel / kTieAbs: the ambiguity window shared with the host resolver

#include <algorithm>
#include <atomic>

namespace dto {

// =====================================================================================================
// K2: critical-overlap tables
// =====================================================================================================
// pass 0: write the kcrit tables and, per cell, kbase = kcrit_1 and count = #k tabulated (k from kcrit_1 up to, not
//         including, kcrit_levels: the range where the screen can pass a cell at a tabulated level)
// pass 1: fill lptab[offset + (k - kbase)] = log p(k) for that range (offsets = exclusive scan of the counts)
__global__ void __launch_bounds__(256) build_kcrit_kernel(const Problem P, uint16_t *__restrict__ kcrit, int pass,
                                                          uint32_t *__restrict__ counts, uint2 *__restrict__ meta,
                                                          double *__restrict__ lptab) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= P.T1 * P.T2) return;
    const int i = cell / P.T2, j = cell % P.T2;
    const uint64_t N = P.N;
    const uint64_t K = P.c1[i], n = P.c2[j];
    const uint64_t lower = (K + n > N) ? (K + n - N) : 0;
    const uint64_t upper = K < n ? K : n;
    if (lower + 1 > upper) {  // no valid k: table stays 0xFFFF
        if (pass == 0) counts[cell] = 0;
        return;
    }
    uint32_t tab_lo = 0, tab_n = 0;  // mode 1: tabulated range
    if (pass == 1) {
        const uint2 m = meta[cell];
        tab_lo = m.y & 0xFFFFu;
        tab_n = m.y >> 16;
        if (tab_n == 0) return;
    }
    const uint32_t never = P.never;  // "no overlap passes": 0xFFFF, or 0x7FFF when the packed 15-bit screen is in use
    uint32_t kc1 = never, kcL = never;  // pass 0: kcrit at level 1 and at the deepest level
    const size_t col = kcrit_col(j, P.CH);
    const size_t level_stride = (size_t)P.T1 * P.T2pad;
    uint16_t *dst = kcrit + (size_t)i * P.T2pad + col;
    if (pass == 0) dst[0] = (uint16_t)(lower + 1);  // level 0: every cell the reference does not short-circuit to p = 1

    const double rowA = P.rowA[i], colB = P.colB[j];
    auto lpmf = [&](uint64_t x) { return log_pmf(P, rowA, colB, (uint32_t)K, (uint32_t)n, (uint32_t)x); };
    const uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    uint64_t xs = mode > lower + 1 ? mode : lower + 1;
    if (xs > upper) xs = upper;
    int l = P.levels;
    if (lpmf(xs) < -80.0) {  // the whole valid range is negligible (support starts far above the mode)
        if (pass == 0) {
            for (; l >= 1; --l) dst[l * level_stride] = (uint16_t)(lower + 1);
            counts[cell] = 0;
            meta[cell] = make_uint2(0u, (uint32_t)(lower + 1));
        }
        return;
    }
    uint64_t x_end = upper;
    if (lpmf(upper) < -80.0) {
        uint64_t lo = xs, hi = upper;  // lpmf(lo) >= -80 > lpmf(hi), lpmf decreasing above the mode
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (lpmf(mid) >= -80.0) lo = mid;
            else hi = mid;
        }
        x_end = hi;
    }
    uint64_t k = x_end;
    double t = exp(lpmf(k));
    double S = 0.0;
    for (;;) {
        S += t;  // p(k) up to a tail below e^-80
        if (pass == 1) {
            if (k >= tab_lo && k < tab_lo + tab_n) lptab[meta[cell].x + (uint32_t)(k - tab_lo)] = log(S);
            if (k <= tab_lo) break;
        } else {
            while (l >= 1 && S > exp(P.level_log[l]) * (1.0 + P.tab_slack)) {
                const uint16_t v = (k + 1 <= upper) ? (uint16_t)(k + 1) : (uint16_t)never;
                dst[l * level_stride] = v;
                if (l == P.levels) kcL = v;
                if (l == 1) kc1 = v;
                --l;
            }
            if (l == 0) break;
            if (k == lower + 1) {
                for (; l >= 1; --l) {
                    dst[l * level_stride] = (uint16_t)(lower + 1);
                    if (l == P.levels) kcL = (uint16_t)(lower + 1);
                    if (l == 1) kc1 = (uint16_t)(lower + 1);
                }
                break;
            }
        }
        // pmf(k-1) = pmf(k) * k (N-K-n+k) / ((K-k+1)(n-k+1))
        t *= ((double)k * (double)(N - K - n + k)) / ((double)(K - k + 1) * (double)(n - k + 1));
        --k;
    }
    if (pass == 0) {
        // tabulate k in [kc1, min(kcL, upper + 1)): cells the screen can pass at levels 1 .. levels-1; deeper (k >= kcL)
        // or shallower (k < kc1, only reachable at level 0) cells take the closed-form / recurrence path in the scan
        uint32_t cnt = 0;
        if (P.levels >= 2 && kc1 != never) {
            const uint32_t hi = (kcL == never) ? (uint32_t)upper + 1u : (uint32_t)kcL;
            cnt = hi > kc1 ? hi - kc1 : 0u;
        }
        counts[cell] = cnt;
        meta[cell] = make_uint2(0u, (uint32_t)kc1 | (cnt << 16));
    }
}

__global__ void fill_u16_kernel(uint16_t *__restrict__ dst, size_t n, uint16_t v) {
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) dst[x] = v;
}

// exclusive scan of counts[0..n) in place (one CTA of 1024 threads: n is a few hundred thousand words, once per problem);
// *total_out = the sum, saturated at 0xFFFFFFFF
__global__ void __launch_bounds__(1024) exclusive_scan_counts_kernel(uint32_t *__restrict__ counts, int n,
                                                                     unsigned long long *__restrict__ total_out) {
    __shared__ unsigned long long wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int b = min(tid * per, n), e = min(b + per, n);
    unsigned long long sum = 0;
    for (int x = b; x < e; ++x) sum += counts[x];
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned long long v = wsum[lane], s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(kFull, s, o);
            if (lane >= o) s += u;
        }
        wsum[lane] = s - v;
        if (lane == 31) *total_out = s;
    }
    __syncthreads();
    unsigned long long run = wsum[w] + inc - sum;
    for (int x = b; x < e; ++x) {
        const uint32_t v = counts[x];
        counts[x] = (uint32_t)(run > 0xFFFFFFFFull ? 0xFFFFFFFFull : run);
        run += v;
    }
}

__global__ void set_meta_offsets_kernel(int cells, const uint32_t *__restrict__ offsets, uint2 *__restrict__ meta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) meta[c].x = offsets[c];
}

// =====================================================================================================
// K0: exactly-uniform random pairing by sorting Philox keys in shared memory.
//   Sort key of element e = (key16(e), sec32(e), e): key16 = one 16-bit half of a Philox4x32-10 word (stream s, one call
//   per 8 elements), sec32 = a word of a second Philox stream (s + 8, one call per 4 elements), evaluated only where
//   key16 does not decide.  i.i.d. keys + index tie-break = a uniform permutation (up to ~n^2 / 2^48).
//   Pass 1: bucket = top B bits of key16 (B ~ log2 n - 1, ~1.2 elements per bucket); ONE shared-memory atomic per
//   element both counts the bucket and hands the element its arrival slot; (key16, arrival) is kept per element so no
//   later pass regenerates the primary keys.  Pass 2: exclusive scan of the counts.  Then either every element ranks
//   itself inside its bucket (exact variant), or only the elements of buckets that straddle a threshold-row boundary do
//   (row-wise variant, the hot path).
// =====================================================================================================
struct SortShared {
    uint32_t *cnt;       // [NB+1]  bucket counts, then offsets (exclusive scan); row-wise variant: bit 31 = "bucket
                         //         straddles a row boundary"
    uint32_t *ba;        // [8 * ceil(n/8)] (key16 << 16) | arrival slot inside the bucket, at ba_index(e); the row-wise
                         //         variant later replaces the entries of boundary-bucket elements by their tie-break word
    uint32_t *scan_tmp;  // [kScanTmp]: warp partials of the block scan (4 slices x 32), [kListCtr] = boundary-list counter
    uint32_t *list_s;    // [list_cap] (shared) boundary-bucket elements: bucket << 16 | element, later position << 16 | element
    uint32_t *list_g;    // [n] (global scratch) the same list beyond list_cap entries
    uint32_t list_cap;
    uint32_t *words;     // [n] (global scratch, exact variant) element at each position, bucket-contiguous
    uint32_t *tie;       // [n] (global scratch, exact variant) tie-break word of every element
};

constexpr uint32_t kPosMask = 0x7FFFFFFFu;
constexpr int kListCtr = 256, kScanTmp = 257;  // block-scan partials: up to 8 slices x 32 warps; then the list counter

__device__ __forceinline__ void philox_keys(uint32_t (&out)[4], uint64_t seed, uint64_t perm_id, uint32_t stream,
                                            uint32_t block) {
    out[0] = block;
    out[1] = stream;
    out[2] = (uint32_t)perm_id;
    out[3] = (uint32_t)(perm_id >> 32);
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint32_t secondary_key(uint64_t seed, uint64_t perm_id, uint32_t stream, uint32_t e) {
    uint32_t c[4];
    philox_keys(c, seed, perm_id, stream + 8u, e >> 2);
    return c[e & 3u];
}

// Order inside a bucket = (tie-break word, element index); the word = the key16 bits below the bucket bits, followed by
// the top 16 + B bits of the secondary key.  Overall order = (key16, sec32 >> (16 - B), index).
__device__ __forceinline__ uint32_t tie_word(uint32_t key16, uint32_t sec32, int B) {
    return ((key16 & ((1u << (16 - B)) - 1u)) << (16 + B)) | (sec32 >> (16 - B));
}

// (key16 | arrival) words are stored so that the two 128-bit accesses of the thread owning key block c8 = e / 8 are
// contiguous across threads (conflict-free): half h = (e >> 2) & 1 of block c8 lives at uint4 index h * n8 + c8.
__device__ __forceinline__ uint32_t ba_index(uint32_t e, uint32_t n8) {
    return ((((e >> 2) & 1u) * n8 + (e >> 3)) << 2) | (e & 3u);
}

// exclusive scan of a[0..len) in place, a[len] = total; blockDim.x threads.  V > 0: every thread owns a contiguous run
// of 4 * V counters held in registers (128-bit shared loads/stores, len == 4 * V * blockDim.x); V == 0: plain serial runs.
template <int V>
__device__ void block_exclusive_scan_t(uint32_t *a, uint32_t len, uint32_t *tmp) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t lane = tid & 31, w = tid >> 5;
    const uint32_t per = V > 0 ? 4u * V : (len + nt - 1) / nt;
    const uint32_t b = tid * per < len ? tid * per : len, e = (b + per < len) ? b + per : len;
    uint4 r[V > 0 ? V : 1];
    uint32_t sum = 0;
    if (V > 0) {
        const uint4 *a4 = reinterpret_cast<const uint4 *>(a + b);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            r[q] = a4[q];
            sum += r[q].x + r[q].y + r[q].z + r[q].w;
        }
    } else {
        for (uint32_t x = b; x < e; ++x) sum += a[x];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) tmp[w] = inc;
    __syncthreads();
    // every warp scans the (at most 32) warp totals itself: one barrier less than handing the job to warp 0
    uint32_t wv = (lane < (nt >> 5)) ? tmp[lane] : 0, ws = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(kFull, ws, o);
        if (lane >= (uint32_t)o) ws += u;
    }
    const uint32_t woff = __shfl_sync(kFull, ws - wv, (int)w);  // exclusive offset of this warp
    uint32_t run = woff + inc - sum;
    if (V > 0) {
        uint4 *a4 = reinterpret_cast<uint4 *>(a + b);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            uint4 o;
            o.x = run;
            o.y = o.x + r[q].x;
            o.z = o.y + r[q].y;
            o.w = o.z + r[q].z;
            run = o.w + r[q].w;
            a4[q] = o;
        }
    } else {
        for (uint32_t x = b; x < e; ++x) {
            const uint32_t v = a[x];
            a[x] = run;
            run += v;
        }
    }
    if (tid == nt - 1) a[len] = run;  // the last thread's run ends at len (or is empty): run == total
    __syncthreads();
}

// The same scan for len == 4 * V * blockDim.x with a bank-conflict-free access pattern: thread t owns the 128-bit vectors
// t, t + nt, .., t + (V-1) nt (consecutive threads touch consecutive vectors), i.e. V slices of 4 nt counters each, scanned
// as V simultaneous block scans whose totals chain from slice to slice.  (Contiguous ownership, 16 V bytes per thread,
// makes every 128-bit access a 4-way conflict at V = 4: a quarter of all shared-memory wavefronts of the pairing kernel.)
// tmp: V * 32 + 1 words.
template <int V>
__device__ void block_exclusive_scan_sliced(uint32_t *a, uint32_t *tmp) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    uint4 *a4 = reinterpret_cast<uint4 *>(a);
    uint4 r[V];
    uint32_t sum[V], inc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        r[q] = a4[(uint32_t)q * nt + tid];
        sum[q] = r[q].x + r[q].y + r[q].z + r[q].w;
        inc[q] = sum[q];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const uint32_t v = __shfl_up_sync(kFull, inc[q], o);
            if (lane >= (uint32_t)o) inc[q] += v;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < V; ++q) tmp[q * 32 + w] = inc[q];
    }
    __syncthreads();
    // every warp scans the V x nw warp totals itself (slice-major order): lane l holds warp l's total of every slice
    uint32_t wv[V], ws[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        wv[q] = lane < nw ? tmp[q * 32 + lane] : 0u;
        ws[q] = wv[q];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const uint32_t u = __shfl_up_sync(kFull, ws[q], o);
            if (lane >= (uint32_t)o) ws[q] += u;
        }
    }
    uint32_t slice_base = 0;
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const uint32_t woff = __shfl_sync(kFull, ws[q] - wv[q], (int)w);  // exclusive offset of this warp inside slice q
        const uint32_t tot = __shfl_sync(kFull, ws[q], 31);
        uint32_t run = slice_base + woff + inc[q] - sum[q];
        uint4 o4;
        o4.x = run;
        o4.y = o4.x + r[q].x;
        o4.z = o4.y + r[q].y;
        o4.w = o4.z + r[q].z;
        a4[(uint32_t)q * nt + tid] = o4;
        slice_base += tot;
    }
    if (tid == 0) a[4u * V * nt] = slice_base;  // total
    __syncthreads();
}

__device__ void block_exclusive_scan(uint32_t *a, uint32_t len, uint32_t *tmp) {
    const uint32_t nt = blockDim.x;
    if (len == 32 * nt) block_exclusive_scan_sliced<8>(a, tmp);
    else if (len == 16 * nt) block_exclusive_scan_sliced<4>(a, tmp);
    else if (len == 8 * nt) block_exclusive_scan_sliced<2>(a, tmp);
    else if (len == 4 * nt) block_exclusive_scan_sliced<1>(a, tmp);
    else block_exclusive_scan_t<0>(a, len, tmp);
}

// Pass 1 + scan, shared by both variants: on return cnt[b] = first position of bucket b (cnt[NB] = n) and ba holds
// (key16 << 16 | arrival) of every element.
__device__ __forceinline__ void block_zero_counters(const SortShared &S, int B) {
    for (uint32_t x = threadIdx.x; x <= (1u << B); x += blockDim.x) S.cnt[x] = 0;
    if (threadIdx.x == 0) S.scan_tmp[kListCtr] = 0;
}

// pre_zeroed: the caller cleared the counters in an earlier phase that a barrier already closed
__device__ void block_bucket_keys(const SortShared &S, uint32_t n, int B, uint64_t seed, uint64_t perm_id, uint32_t stream,
                                  bool pre_zeroed = false) {
    const uint32_t NB = 1u << B;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    if (!pre_zeroed) {
        block_zero_counters(S, B);
        __syncthreads();
    }
    const uint32_t n8 = (n + 7) >> 3;
    uint4 *ba4 = reinterpret_cast<uint4 *>(S.ba);
    for (uint32_t c = tid; c < n8; c += nt) {
        uint32_t key[4];
        philox_keys(key, seed, perm_id, stream, c);
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t e = c * 8 + q;
            const uint32_t k16 = (q & 1) ? (key[q >> 1] >> 16) : (key[q >> 1] & 0xFFFFu);
            v[q] = k16 << 16;
            if (e < n) v[q] |= atomicAdd(&S.cnt[k16 >> (16 - B)], 1u);
        }
        ba4[c] = make_uint4(v[0], v[1], v[2], v[3]);
        ba4[n8 + c] = make_uint4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    block_exclusive_scan(S.cnt, NB, S.scan_tmp);
}

// Exact variant: ranks the n elements of one Philox stream; calls emit(e, f): element e has the f-th smallest key.
template <typename Emit>
__device__ void block_rank_by_random_keys(const SortShared &S, uint32_t n, int B, uint64_t seed, uint64_t perm_id,
                                          uint32_t stream, Emit emit) {
    block_bucket_keys(S, n, B, seed, perm_id, stream);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t n8 = (n + 7) >> 3, nblk = (n + 3) >> 2;
    for (uint32_t c = tid; c < nblk; c += nt) {
        uint32_t sk[4];
        philox_keys(sk, seed, perm_id, stream + 8u, c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t e = c * 4 + q;
            if (e < n) {
                const uint32_t v = S.ba[ba_index(e, n8)];
                S.words[S.cnt[v >> (32 - B)] + (v & 0xFFFFu)] = e;
                S.tie[e] = tie_word(v >> 16, sk[q], B);
            }
        }
    }
    __syncthreads();
    for (uint32_t e = tid; e < n; e += nt) {
        const uint32_t v = S.ba[ba_index(e, n8)];
        const uint32_t b = v >> (32 - B);
        const uint32_t lo = S.cnt[b], hi = S.cnt[b + 1];
        uint32_t rank = 0;
        if (hi - lo > 1) {
            const uint32_t mine = S.tie[e];
            for (uint32_t y = lo; y < hi; ++y) {
                const uint32_t ey = S.words[y];
                const uint32_t ty = S.tie[ey];
                rank += (ty < mine || (ty == mine && ey < e)) ? 1u : 0u;
            }
        }
        emit(e, lo + rank);
    }
    __syncthreads();
}

// ---- explicit shared-space accesses for the hot loops of the row-wise variant: the pointers of SortShared are generic
// ---- (they may also point into the global scratch), and a generic atomic / load costs an address-space conversion each
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t atoms_inc(uint32_t addr) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg128_plain(const uint32_t *p) {  // data written by this kernel: no read-only path
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg128(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

// bucket counters (and the boundary-list counter) back to zero, 128 bits per store
__device__ __forceinline__ void block_zero_counters_fast(const SortShared &S, int B) {
    const uint32_t NB = 1u << B;
    if (NB >= 4) {
        const uint32_t base = smem_addr(S.cnt);
        for (uint32_t x = threadIdx.x; x < NB / 4; x += blockDim.x) sts128(base + x * 16, 0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) S.cnt[NB] = 0;
    } else {
        for (uint32_t x = threadIdx.x; x <= NB; x += blockDim.x) S.cnt[x] = 0;
    }
    if (threadIdx.x == 0) S.scan_tmp[kListCtr] = 0;
}

// Pass 1 of the row-wise variant (S.ba in shared memory): same result as block_bucket_keys, branch-free per element.
// Blocks of 8 elements that lie entirely below n take the lean path; the one partial block (n % 8 != 0) is checked.
template <bool BA_SHARED>
__device__ __forceinline__ void block_bucket_keys_lean(const SortShared &S, uint32_t n, int B, uint64_t seed,
                                                       uint64_t perm_id, uint32_t stream) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t n8 = (n + 7) >> 3, full = n >> 3;
    const uint32_t cnt_a = smem_addr(S.cnt), ba_a = BA_SHARED ? smem_addr(S.ba) : 0u;
    // (key16 | arrival) words: in shared memory, or -- so that two CTAs fit one SM at N = 20 000 -- in the CTA's global
    // scratch (L2-resident; written here and read by the placement with coalesced 128-bit accesses)
    auto put_ba = [&](uint32_t vec, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        if constexpr (BA_SHARED) sts128(ba_a + vec * 16, a, b, c, d);
        else stg128(S.ba + (size_t)vec * 4, a, b, c, d);
    };
    const uint32_t sh = 32u - (uint32_t)B;  // bucket of the key in the HIGH half of a Philox word: w >> sh
    for (uint32_t c = tid; c < full; c += nt) {
        uint32_t key[4];
        philox_keys(key, seed, perm_id, stream, c);
        uint32_t v[8];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t w = key[h];  // low half = element 8c + 2h, high half = element 8c + 2h + 1
            const uint32_t wl = w << 16;
            const uint32_t a0 = atoms_inc(cnt_a + ((wl >> sh) << 2));
            const uint32_t a1 = atoms_inc(cnt_a + ((w >> sh) << 2));
            v[2 * h] = wl | a0;
            v[2 * h + 1] = (w & 0xFFFF0000u) | a1;
        }
        put_ba(c, v[0], v[1], v[2], v[3]);
        put_ba(n8 + c, v[4], v[5], v[6], v[7]);
    }
    if (full < n8 && tid == (full % nt)) {  // the partial block
        const uint32_t c = full;
        uint32_t key[4];
        philox_keys(key, seed, perm_id, stream, c);
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t e = c * 8 + q;
            const uint32_t k16 = (q & 1) ? (key[q >> 1] >> 16) : (key[q >> 1] & 0xFFFFu);
            v[q] = k16 << 16;
            if (e < n) v[q] |= atoms_inc(cnt_a + ((k16 >> (16 - B)) << 2));
        }
        put_ba(c, v[0], v[1], v[2], v[3]);
        put_ba(n8 + c, v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    block_exclusive_scan(S.cnt, 1u << B, S.scan_tmp);
}

// Row-wise variant, used when only the ROW of every list-1 position matters (identical gene sets, no export): positions
// inside one threshold row are interchangeable for the overlap grid, so an element whose whole bucket lies inside one
// row takes position off[bucket] + arrival without being ranked.  Buckets that contain a row boundary strictly inside
// are found from the boundaries' side (one binary search over the bucket offsets per threshold, `bounds` = Problem::c1);
// only their elements (~5-8 %) get a secondary key and an exact rank, entirely in shared memory: the staged row itself
// holds the member list of such a bucket until the ranks are known.  The result is the exact variant's permutation up
// to within-row order; the records it leads to are identical.
//
// The placement loop touches every element and is therefore kept minimal (per element: bucket offset, position, one
// select, one 16-bit store); everything about boundary buckets happens on the side: the thread that owns the FIRST
// boundary inside such a bucket lists its members afterwards (they sit in the staged row), and the ~6 % of elements on
// that list are then ranked by all threads.
template <bool BA_SHARED, int NT>
__device__ void block_place_rowwise(const SortShared &S, const Problem &P, const uint32_t *bounds, uint32_t n, int B,
                                    uint64_t seed, uint64_t perm_id, uint32_t stream, uint16_t *stage) {
    constexpr int NBND = 2048 / NT;  // row boundaries per thread (T1 <= 2048)
    block_bucket_keys_lean<BA_SHARED>(S, n, B, seed, perm_id, stream);
    const uint32_t NB = 1u << B;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t lane = tid & 31u;
    // the boundary list: first list_cap entries in shared memory, the rest in the global scratch (two explicit branches
    // so that each side is a plain shared / global access, not a generic one)
    auto lst_put = [&](uint32_t x, uint32_t v) {
        if (x < S.list_cap) S.list_s[x] = v;
        else S.list_g[x] = v;
    };
    auto lst_get = [&](uint32_t x) -> uint32_t { return x < S.list_cap ? S.list_s[x] : S.list_g[x]; };
    // The bucket a row boundary p cuts (off[b] < p < off[b+1]), per boundary; T1 <= 2048 = 2 boundaries per thread.  The
    // thread whose boundary is the FIRST one inside a bucket (the previous boundary lies at or before its start) owns it.
    uint32_t own[NBND], own_m[NBND], own_lo[NBND];
#pragma unroll
    for (int r = 0; r < NBND; ++r) {
        own[r] = 0xFFFFFFFFu, own_m[r] = 0u, own_lo[r] = 0u;
        const uint32_t x = tid + (uint32_t)r * nt;
        if (x < (uint32_t)P.T1) {
            const uint32_t p = bounds[x];
            if (p > 0 && p < n) {
                uint32_t lo = 0, hi = NB;  // largest b with off[b] <= p: the (non-empty) bucket holding position p
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (S.cnt[mid] <= p) lo = mid;
                    else hi = mid;
                }
                const uint32_t start = S.cnt[lo];
                if (start < p && (x == 0 || bounds[x - 1] <= start)) {
                    own[r] = lo;
                    own_lo[r] = start;
                    own_m[r] = S.cnt[lo + 1] - start;
                }
            }
        }
    }
    __syncthreads();
    // owners flag their bucket (bit 31 of its offset; one owner per bucket: a plain store) and reserve one list entry per
    // member, (bucket << 16 | member index) -- the member's element id is only known after the placement, which leaves
    // it in the staged row.  One warp-aggregated reservation per warp that owns anything.
#pragma unroll
    for (int r = 0; r < NBND; ++r) {
        if ((uint32_t)r * nt >= (uint32_t)P.T1) break;                    // block-uniform
        if ((tid & ~31u) + (uint32_t)r * nt >= (uint32_t)P.T1) continue;  // warp-uniform: no boundary in this warp
        const uint32_t m = own_m[r];
        if (m) S.cnt[own[r]] = own_lo[r] | 0x80000000u;
        uint32_t inc = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        uint32_t base = 0;
        if (lane == 31 && inc) base = atomicAdd(&S.scan_tmp[kListCtr], inc);
        base = __shfl_sync(kFull, base, 31) + inc - m;
        for (uint32_t y = 0; y < m; ++y) lst_put(base + y, (own[r] << 16) | y);
    }
    // the previous permutation's staged row is still being read by its bulk copy-out (issued by thread 0 after the last
    // barrier of that permutation): the placement below is the first phase to overwrite it
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    const uint32_t n8 = (n + 7) >> 3;
    const uint32_t steps = 2 * n8;
    {
        // placement, 4 elements (one 128-bit word of ba, one 8-byte word of partner slots) per step.  An element of a
        // boundary bucket stores its own index (the staged row doubles as the member list of such a bucket; it holds
        // >= n entries), any other element its partner slot.  Positions past the last threshold are overwritten with
        // kNoSlot after the ranking.  Only the two steps of the partial key block (n % 8 != 0) check e < n.
        const uint32_t cnt_a = smem_addr(S.cnt), st_a = smem_addr(stage);
        const uint32_t ba_a = BA_SHARED ? smem_addr(S.ba) : 0u;
        const uint32_t sh = 32u - (uint32_t)B;
        const bool tail = (n & 7u) != 0u;
        for (uint32_t u = tid; u < steps; u += nt) {
            const uint32_t e0 = u < n8 ? u * 8 : (u - n8) * 8 + 4;
            uint4 va;
            if constexpr (BA_SHARED) va = lds128(ba_a + u * 16);
            else va = ldg128_plain(S.ba + (size_t)u * 4);
            const uint2 dsv = __ldg(reinterpret_cast<const uint2 *>(P.dslot2 + e0));  // dslot2 is padded to a multiple of 8
            const uint32_t v[4] = {va.x, va.y, va.z, va.w};
            const uint32_t ds[4] = {dsv.x & 0xFFFFu, dsv.x >> 16, dsv.y & 0xFFFFu, dsv.y >> 16};
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = lds32(cnt_a + ((v[q] >> sh) << 2));
            if (!tail || (u != n8 - 1 && u != steps - 1)) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t pos = (w[q] & kPosMask) + (v[q] & 0xFFFFu);
                    sts16(st_a + 2 * pos, (int32_t)w[q] < 0 ? e0 + q : ds[q]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t pos = (w[q] & kPosMask) + (v[q] & 0xFFFFu);
                    if (e0 + q < n) sts16(st_a + 2 * pos, (int32_t)w[q] < 0 ? e0 + q : ds[q]);
                }
            }
        }
    }
    __syncthreads();
    const uint32_t n_list = S.scan_tmp[kListCtr];
    for (uint32_t x = tid; x < n_list; x += nt) {
        const uint32_t ent = lst_get(x);
        const uint32_t b = ent >> 16;
        const uint32_t e = stage[(S.cnt[b] & kPosMask) + (ent & 0xFFFFu)];  // the member's element id
        uint32_t &slot = S.ba[ba_index(e, n8)];
        slot = tie_word(slot >> 16, secondary_key(seed, perm_id, stream, e), B);
        lst_put(x, (b << 16) | e);
    }
    __syncthreads();
    for (uint32_t x = tid; x < n_list; x += nt) {
        const uint32_t ent = lst_get(x);
        const uint32_t e = ent & 0xFFFFu, b = ent >> 16;
        const uint32_t lo = S.cnt[b] & kPosMask, hi = S.cnt[b + 1] & kPosMask;
        const uint32_t mine = S.ba[ba_index(e, n8)];
        uint32_t rank = 0;
        for (uint32_t y = lo; y < hi; ++y) {
            const uint32_t ey = stage[y];
            const uint32_t ty = S.ba[ba_index(ey, n8)];
            rank += (ty < mine || (ty == mine && ey < e)) ? 1u : 0u;
        }
        lst_put(x, ((lo + rank) << 16) | e);
    }
    __syncthreads();
    for (uint32_t x = tid; x < n_list; x += nt) {
        const uint32_t ent = lst_get(x);
        const uint32_t f = ent >> 16;
        if (f < P.n1_eff) stage[f] = P.dslot2[ent & 0xFFFFu];
    }
    // positions past the last threshold carry no partner (disjoint from the writes above)
    for (uint32_t x = P.n1_eff + tid; x < P.pb_stride; x += nt) stage[x] = kNoSlot;
    __syncthreads();
}

// NT threads per CTA: 1024 (one CTA per SM) or 512 (two CTAs per SM, each on its own permutation, so that the block
// barriers and shared-memory round trips of one overlap with the other's work)
template <bool BA_IN_SCRATCH, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) sigma_sort_kernel(const __grid_constant__ Problem P, uint64_t seed0,
                                                                   const uint64_t *__restrict__ seeds, uint32_t seg,
                                                                   uint64_t first_id, int n_tasks, int B1, int B2,
                                                                   uint16_t *__restrict__ pb,
                                                                   uint32_t *__restrict__ pairing_out,
                                                                   uint32_t *__restrict__ scratch_base,
                                                                   uint32_t list_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t nmax = P.n1 > P.n2 ? P.n1 : P.n2;
    const uint32_t nmax8 = (nmax + 7) & ~7u;
    const int Bmax = B1 > B2 ? B1 : B2;
    // shared: cnt[NB+1] | scan_tmp[kScanTmp] | pad | stage[max(pb_stride, nmax8)] u16 | bounds[T1] u32 | order2[n_common] u16 |
    //         ba[nmax8] u32 (unless in scratch) | list_s[list_cap] u32
    // per-CTA global scratch (L2-resident): words[nmax8] | tie[nmax8] | list_g[nmax8] | ba[nmax8] (when it does not fit)
    SortShared S;
    S.cnt = reinterpret_cast<uint32_t *>(smem_raw);
    S.scan_tmp = S.cnt + (1u << Bmax) + 1;
    size_t off = (((size_t)(1u << Bmax) + 1 + kScanTmp) * 4 + 15) & ~(size_t)15;
    uint16_t *stage = reinterpret_cast<uint16_t *>(smem_raw + off);
    off += (size_t)(P.pb_stride > nmax8 ? P.pb_stride : nmax8) * 2;  // both are multiples of 8 entries: 16 B aligned
    uint32_t *bounds = reinterpret_cast<uint32_t *>(smem_raw + off);
    off += (((size_t)P.T1 * 4) + 15) & ~(size_t)15;
    const bool identical = (P.n_common == P.n1 && P.n_common == P.n2);
    uint16_t *order2 = reinterpret_cast<uint16_t *>(smem_raw + off);
    if (!identical) off += (((size_t)P.n_common * 2) + 15) & ~(size_t)15;
    uint32_t *scratch = scratch_base + (size_t)blockIdx.x * 4 * nmax8;
    S.words = scratch;
    S.tie = scratch + nmax8;
    S.list_g = scratch + 2 * (size_t)nmax8;
    if constexpr (BA_IN_SCRATCH) {  // a template parameter so that the shared-memory case compiles to LDS/STS, not generic
        S.ba = scratch + 3 * (size_t)nmax8;
    } else {
        S.ba = reinterpret_cast<uint32_t *>(smem_raw + off);
        off += (size_t)nmax8 * 4;
    }
    S.list_s = reinterpret_cast<uint32_t *>(smem_raw + off);
    S.list_cap = list_cap;
    for (uint32_t x = threadIdx.x; x < (uint32_t)P.T1; x += blockDim.x) bounds[x] = P.c1[x];
    __syncthreads();
    const bool rowwise = identical && pairing_out == nullptr;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    if (rowwise) {
        block_zero_counters_fast(S, B2);
        __syncthreads();
    }

    for (int t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        // batched list pairs: `seg` consecutive tasks belong to one pair, which has its own seed and ids first_id ..
        const uint64_t seed = seeds ? seeds[(uint32_t)t / seg] : seed0;
        const uint64_t perm_id = first_id + (uint64_t)(seeds ? (uint32_t)t % seg : (uint32_t)t);
        uint16_t *dst = pb + (size_t)t * P.pb_stride;
        if (rowwise) {
            block_place_rowwise<!BA_IN_SCRATCH, NT>(S, P, bounds, P.n2, B2, seed, perm_id, 0u, stage);
        } else if (identical) {
            // element = list-2 position e, rank f = the list-1 position it is paired with
            block_rank_by_random_keys(S, P.n2, B2, seed, perm_id, 0u, [&](uint32_t e, uint32_t f) {
                if (f < P.n1_eff) stage[f] = P.dslot2[e];
                if (pairing_out) pairing_out[(size_t)t * P.n1 + f] = e;
            });
            for (uint32_t x = P.n1_eff + tid; x < P.pb_stride; x += nt) stage[x] = kNoSlot;
            __syncthreads();
        } else {
            // uniform random partial injection: the n_common lowest-keyed positions of each list, matched by rank
            block_rank_by_random_keys(S, P.n2, B2, seed, perm_id, 1u, [&](uint32_t e, uint32_t f) {
                if (f < P.n_common) order2[f] = (uint16_t)e;
            });
            // the second sort's emit writes stage[e] for every list-1 position e < n1_eff
            block_rank_by_random_keys(S, P.n1, B1, seed, perm_id, 0u, [&](uint32_t e, uint32_t f) {
                uint32_t partner = 0xFFFFFFFFu;
                if (f < P.n_common) partner = order2[f];
                if (e < P.n1_eff) stage[e] = partner != 0xFFFFFFFFu ? P.dslot2[partner] : kNoSlot;
                if (pairing_out) pairing_out[(size_t)t * P.n1 + e] = partner;
            });
            for (uint32_t x = P.n1_eff + tid; x < P.pb_stride; x += nt) stage[x] = kNoSlot;
            __syncthreads();
        }
        // copy-out (pb_stride is a multiple of 256, rows are 512 B aligned); the row-wise variant clears its bucket
        // counters for the next permutation in the same phase
        if (rowwise) {
            // one 1-D bulk copy (TMA) of the whole row, shared -> global, issued by one thread and overlapped with the
            // next permutation's key pass, block scan and boundary search, none of which touch the staged row
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the row was written through the generic proxy
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(stage)),
                             "r"(P.pb_stride * 2u)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            block_zero_counters_fast(S, B2);
        } else {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(stage);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            for (uint32_t x = tid; x < P.pb_stride / 8; x += nt) d4[x] = s4[x];
        }
        __syncthreads();
    }
    if (rowwise && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the CTA's shared memory outlives the copy
}

// =====================================================================================================
// K0': pairing from host-supplied indices.  Position j of (permuted) list 1 holds the gene of slot perm1[j];
// that gene sits at slot slot2_of_1[.] of list 2, which the permuted list 2 shows at position inv2[slot].
// =====================================================================================================
__global__ void invert_perm_kernel(const uint32_t *__restrict__ perm, uint32_t n, int n_tasks,
                                   uint32_t *__restrict__ inv, int *__restrict__ err) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n_tasks) return;
    const size_t t = idx / n;
    const uint32_t v = perm[idx];
    if (v >= n) {
        atomicExch(err, 1);
        return;
    }
    inv[t * n + v] = (uint32_t)(idx - t * n);
}

__global__ void check_perm_kernel(const uint32_t *__restrict__ perm, uint32_t n, int n_tasks,
                                  const uint32_t *__restrict__ inv, int *__restrict__ err) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n_tasks) return;
    const size_t t = idx / n;
    const uint32_t v = perm[idx];
    if (v >= n || inv[t * n + v] != (uint32_t)(idx - t * n)) atomicExch(err, 1);  // duplicate entry
}

// slot2_maps != nullptr: task t uses its own gene map slot2_maps[t * n1 ..] (batched list pairs that share one rank
// structure but not their gene order); otherwise every task uses P.slot2_of_1.
__global__ void compose_pairing_kernel(const Problem P, const uint32_t *__restrict__ perm1,
                                       const uint32_t *__restrict__ inv2, const int32_t *__restrict__ slot2_maps,
                                       int n_tasks, uint16_t *__restrict__ pb) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P.pb_stride * n_tasks) return;
    const size_t t = idx / P.pb_stride;
    const int32_t *__restrict__ slot2_base = slot2_maps ? slot2_maps + t * P.n1 : P.slot2_of_1;
    const uint32_t j = (uint32_t)(idx - t * P.pb_stride);
    uint16_t v = kNoSlot;
    if (j < P.n1_eff) {
        // perm1 / inv2 come from the caller: an invalid row (entry out of range, or a duplicate that left an inverse slot
        // at its 0xFFFFFFFF fill) already raised *err in the validation kernels; here it must merely not index out of bounds
        const uint32_t slot1 = perm1 ? perm1[t * P.n1 + j] : j;
        if (slot1 < P.n1) {
            const int32_t slot2 = slot2_base[slot1];
            if (slot2 >= 0 && (uint32_t)slot2 < P.n2) {
                const uint32_t pos2 = inv2 ? inv2[t * P.n2 + (uint32_t)slot2] : (uint32_t)slot2;
                if (pos2 < P.n2) v = P.dslot2[pos2];
            }
        }
    }
    pb[idx] = v;
}

// =====================================================================================================
// K1: scan kernel
// =====================================================================================================
struct __align__(16) Cand {
    uint32_t ij;  // row << 16 | column
    uint32_t k;   // overlap | kRefinedBit | kEvalBit
    double v;     // log lower bound of p; once kEvalBit is set: the exact p-value
};

template <int CH>
struct ScanLayout {
    static constexpr int CHP = hist_words(CH);   // packed histogram words per lane (dto_device.cuh)
    static constexpr int QCAP = 32 * CH + 32;    // < 32 left-overs + one full row
    static constexpr int CAP = kCandCap;
    static constexpr uint32_t dummy_off = 32u * CHP * 4u;  // byte offset of the sink word that absorbs kNoSlot partners
    static constexpr size_t d_bytes = (size_t)32 * CHP * 4 + 16;  // CHP is a multiple of 4: 16-byte multiple
    static constexpr size_t q_bytes = ((size_t)QCAP * 6 + 15) & ~(size_t)15;  // u32 (row<<16|col) + u16 k per entry
    static constexpr size_t ring_bytes = (size_t)kRing * 2;  // partner-slot staging ring (cp.async), 4 chunks of kChunk
    static constexpr size_t per_warp = d_bytes + q_bytes + 16 + (size_t)CAP * sizeof(Cand) + ring_bytes;
};

// Per-warp state of the rare path (everything behind the screen).  Lives in local memory on purpose: the row loop
// touches none of it, so its registers stay free for the 2 x CH column state.
//
// Exact ties: the reference picks its optimum with `==` / `<` on p-values computed with the HOST libm's exp()
// (optimize_main.rs:73-80).  The device's exp() may differ from it in the last ulp, so the device never decides between
// two cells whose p-values are closer than the ambiguity window (kTieRel / kTieAbs, dto_host_math.hpp) unless their
// (K, n, k) are equal (then both sides compute identical bits and the integer tie-break is exact).  Instead the cells
// inside the window of the running minimum -- the "tie set" -- stay in the candidate buffer, already evaluated, and at the
// end of the task a tie set with more than one distinct (K, n, k) is shipped to the host (TieEntry pool), which
// re-evaluates it in statrs order with the host libm and applies the reference's tie-break (dto_engine.cu).
struct Rare {
    double theta;  // certified: log(min p of the reference) <= theta
    double pmin;   // smallest non-zero exact p seen so far (warp-uniform); +inf = none
    Best zero;     // per lane: best cell on the underflow plateau (reference p == 0.0), integer tie-break only
    uint32_t ncand;
    uint32_t overflow;  // the tie set outgrew the buffer: the task is re-run through the dense path
    uint32_t n_level2, n_eval, n_refine;
    const uint32_t *s_c1;
    uint32_t *Qij;
    uint16_t *Qk;
    uint32_t *qcnt;
    Cand *cand;
    int lane;
};

// drops buffered candidates that can no longer be within kRefineEps of the minimum; evaluated entries (the tie set) are
// only ever filtered by evaluate_buffer
__device__ __noinline__ void compact_cands(Rare &R) {
    const int lane = R.lane;
    Cand *c = R.cand;
    const uint32_t n = R.ncand;
    const double theta = R.theta;
    uint32_t out = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t idx = b + lane;
        Cand e;
        bool keep = false;
        if (idx < n) {
            e = c[idx];
            keep = (e.k & kEvalBit) || (e.v - kEps <= theta);
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) c[out + __popc(bal & lt)] = e;
        out += __popc(bal);
        __syncwarp();
    }
    R.ncand = out;
}

// (3b) refine: tail by the ratio recurrence (no exp, no table walk) -> log p to ~1e-10 for every buffered candidate
//      not refined yet; tightens theta so that only cells within ~1e-8 of the minimum survive
__device__ __noinline__ void refine_buffer(const Problem &P, Rare &R) {
    double th = CUDART_INF;
    for (uint32_t idx = R.lane; idx < R.ncand; idx += 32) {
        Cand c = R.cand[idx];
        if (c.k & kRefinedBit) continue;
        const uint32_t i = c.ij >> 16, j = c.ij & 0xFFFFu;
        const uint32_t K = R.s_c1[i], n = P.c2[j], k = c.k;
        const double s = log_pmf(P, P.rowA[i], P.colB[j], K, n, k);
        double a = (double)(K - k), b = (double)(n - k);
        double cc = (double)k + 1.0, d = (double)(P.N - K - n + k) + 1.0;
        double t = 1.0, S = 1.0;
        // pmf(k+1)/pmf(k) = a b / (cc d); then a, b fall and cc, d rise by one per step
        while (a > 0.0 && b > 0.0) {
            t *= (a * b) / (cc * d);
            S += t;
            if (t < S * 0x1p-53 && (a * b) < (cc * d)) break;
            a -= 1.0;
            b -= 1.0;
            cc += 1.0;
            d += 1.0;
        }
        const double lp = s + log(S);
        c.v = lp - P.refine_eps + kEps;
        c.k |= kRefinedBit;
        R.cand[idx] = c;
        th = fmin(th, lp + P.refine_eps);
        ++R.n_refine;
    }
    R.theta = fmin(R.theta, warp_min(th));
    __syncwarp();
}

// (4) statrs-order FP64 tail for every buffered candidate not evaluated yet, one per lane; tightens theta to the exact
//     minimum seen so far and shrinks the buffer to the tie set of that minimum (entries keep their exact p in `v`)
__device__ __noinline__ void evaluate_buffer(const Problem &P, Rare &R) {
    const int lane = R.lane;
    Cand *c = R.cand;
    const uint32_t n = R.ncand;
    double mn = CUDART_INF;
    bool sawzero = false;
    for (uint32_t idx = lane; idx < n; idx += 32) {
        Cand e = c[idx];
        if (!(e.k & kEvalBit)) {
            const uint32_t i = e.ij >> 16, j = e.ij & 0xFFFFu, k = e.k & kOverlapMask;
            e.v = hypergeom_pvalue_exact(P.lf, P.N, R.s_c1[i], P.c2[j], k);
            e.k = k | kEvalBit | kRefinedBit;
            c[idx] = e;
            ++R.n_eval;
            if (!(e.v > 0.0)) {  // underflow plateau: exactly 0.0 in the reference too -> integer tie-break
                Best z;
                z.p = 0.0;
                z.k = k;
                z.ij = e.ij;
                if (better(z, R.zero)) R.zero = z;
                sawzero = true;
            }
        }
        if (e.v > 0.0) mn = fmin(mn, e.v);
    }
    const double pmin = fmin(R.pmin, warp_min(mn));
    R.pmin = pmin;
    double th = pmin < CUDART_INF ? log(pmin) + kEps : CUDART_INF;
    if (__any_sync(kFull, sawzero)) th = fmin(th, kZeroHi);
    R.theta = fmin(R.theta, th);
    // keep the tie set: evaluated, non-zero, inside the ambiguity window of the minimum (same index <-> lane mapping as
    // the loop above, so every lane re-reads only what it wrote itself)
    uint32_t out = 0;
    const unsigned lt = (1u << lane) - 1u;
    const double lim = pmin * (1.0 + kTieRel) + kTieAbs;
    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t idx = b + lane;
        Cand e;
        bool keep = false;
        if (idx < n) {
            e = c[idx];
            keep = e.v > 0.0 && e.v <= lim;
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) c[out + __popc(bal & lt)] = e;
        out += __popc(bal);
        __syncwarp();
    }
    if (out > 4) {
        // Lists whose ranks have gaps give equal set sizes at consecutive thresholds, hence cells with the same (K, n, k):
        // bit-identical p on host and device alike, so among them only the smallest (row, column) can win -- drop the rest
        // (keeps the tie set at the number of DISTINCT (K, n, k) in the window).  out <= kCandCap = 64: two entries per lane.
        bool kp[2] = {false, false};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t idx = (uint32_t)h * 32 + lane;
            if (idx < out) {
                const Cand e = c[idx];
                const uint32_t K = R.s_c1[e.ij >> 16], nn = P.c2[e.ij & 0xFFFFu];
                bool dominated = false;
                for (uint32_t y = 0; y < out && !dominated; ++y) {
                    const Cand f = c[y];
                    dominated = f.ij < e.ij && f.k == e.k && R.s_c1[f.ij >> 16] == K && P.c2[f.ij & 0xFFFFu] == nn;
                }
                kp[h] = !dominated;
            }
        }
        __syncwarp();
        uint32_t out2 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t idx = (uint32_t)h * 32 + lane;
            Cand e;
            if (idx < out) e = c[idx];
            const unsigned bal = __ballot_sync(kFull, kp[h]);
            __syncwarp();
            if (kp[h]) c[out2 + __popc(bal & lt)] = e;
            out2 += __popc(bal);
            __syncwarp();
        }
        out = out2;
    }
    R.ncand = out;
}

// (3a) drains the queue of cells that passed the critical-overlap screen, 32 at a time: table lookup of log p (or,
//      outside the tabulated range, certified / closed-form bounds), running bound theta, candidate buffer.
//      Returns the screen level theta now allows.
__device__ __noinline__ int drain_queue(const Problem &P, Rare &R, bool flush, int level) {
    const int lane = R.lane;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t qc = *R.qcnt;
    while (qc >= 32 || (flush && qc > 0)) {
        const uint32_t take = qc < 32 ? qc : 32;
        const uint32_t start = qc - take;
        const bool valid = (uint32_t)lane < take;
        Cand e;
        e.ij = 0;
        e.k = 0;
        e.v = CUDART_INF;
        double ub = CUDART_INF;
        bool keep = false;
        if (valid) {
            e.ij = R.Qij[start + lane];
            e.k = R.Qk[start + lane];
            const uint32_t i = e.ij >> 16, j = e.ij & 0xFFFFu;
            const uint32_t K = R.s_c1[i], n = P.c2[j], k = e.k;
            const uint2 meta = P.cellmeta[(size_t)i * P.T2 + j];
            const uint32_t dk = k - (meta.y & 0xFFFFu);
            if (dk < (meta.y >> 16)) {
                // tabulated: log p of this (cell, k) to ~1e-10, no arithmetic at all
                const double lp = P.lptab[meta.x + dk];
                ub = lp + P.refine_eps - kEps;
                e.v = lp - P.refine_eps + kEps;
                e.k |= kRefinedBit;
                keep = true;
            } else if (k < (meta.y & 0xFFFFu)) {
                // below the tabulated range: the table build certified p(kbase - 1) > tau_1, and p falls with k
                e.v = P.level_log[1];
                keep = true;
            } else {
                const double s = log_pmf(P, P.rowA[i], P.colB[j], K, n, k);
                const double a = (double)(K - k), b = (double)(n - k);
                const double c = (double)k + 1.0, d = (double)(P.N - K - n + k) + 1.0;
                const double r1 = (a * b) / (c * d);  // pmf(k+1) / pmf(k)
                if (r1 < 1.0 && s < kZeroLo) {
                    // every tail term underflows: the reference's p is exactly 0.0 -> integer tie-break only
                    Best z;
                    z.p = 0.0;
                    z.k = k;
                    z.ij = e.ij;
                    if (better(z, R.zero)) R.zero = z;
                    ub = kZeroHi - kEps;
                } else {
                    double lb = s;  // p >= pmf(k)
                    if (r1 < 1.0) {
                        ub = s - log1p(-r1) + P.refine_eps;  // ratios fall with k: p <= pmf(k) / (1 - r1)
                        // ratios over the next mm steps are all >= r_mm: p >= pmf * (1 - r^(mm+1)) / (1 - r)
                        double mm = floor(2.0 / (1.0 - r1)) + 1.0;
                        mm = fmin(mm, fmin(a, b));
                        if (mm >= 1.0) {
                            const double rm = ((a - mm + 1.0) * (b - mm + 1.0)) / ((c + mm - 1.0) * (d + mm - 1.0));
                            if (rm > 0.0 && rm < 1.0) lb = s + log((1.0 - exp((mm + 1.0) * log(rm))) / (1.0 - rm));
                        }
                    }
                    e.v = lb - P.refine_eps;  // log pmf itself carries the table's rounding (ulp(lf[N]) for huge populations)
                    keep = true;
                }
            }
        }
        R.theta = fmin(R.theta, warp_min(ub + kEps));
        keep = keep && (e.v - kEps <= R.theta);
        const unsigned bal = __ballot_sync(kFull, keep);
        if (R.ncand + __popc(bal) > (uint32_t)kCandCap) {
            compact_cands(R);
            if (R.ncand + __popc(bal) > (uint32_t)kCandCap) {  // still full: sharpen the survivors' bounds
                refine_buffer(P, R);
                compact_cands(R);
            }
            if (R.ncand + __popc(bal) > (uint32_t)kCandCap) evaluate_buffer(P, R);  // genuinely full of near-ties
            if (R.ncand + __popc(bal) > (uint32_t)kCandCap) {  // a tie set of > 32 cells: hand the task to the dense path
                R.overflow = 1u;
                R.ncand = 0;
            }
            keep = keep && (e.v - kEps <= R.theta);
        }
        const unsigned bal2 = __ballot_sync(kFull, keep);
        if (keep) R.cand[R.ncand + __popc(bal2 & lt)] = e;
        R.ncand += __popc(bal2);
        __syncwarp();
        R.n_level2 += take;
        qc = start;
    }
    __syncwarp();  // every lane has read the old count before lane 0 replaces it
    if (lane == 0) *R.qcnt = qc;
    __syncwarp();
    while (level < P.levels && R.theta <= P.level_log[level + 1]) ++level;
    return level;
}

// (5) end of a permutation: settle what is left, warp-shuffle argmin with the reference tie-break, write the record
__device__ __noinline__ void finish_task(const Problem &P, Rare &R, int task, int level, uint32_t record_flags,
                                         dto_b200_record *__restrict__ out, const ScanOut &O,
                                         unsigned long long *__restrict__ counters, uint32_t *__restrict__ task_stats,
                                         long long t_begin) {
    const int lane = R.lane;
    level = drain_queue(P, R, true, level);
    compact_cands(R);
    refine_buffer(P, R);
    compact_cands(R);
    evaluate_buffer(P, R);
    // the buffer now holds the tie set of the smallest non-zero p; R.zero the best cell of the underflow plateau
    const Best zero = warp_best(R.zero);
    const uint32_t nt = R.ncand;
    Best mine;
    mine.p = CUDART_INF;
    mine.k = 0;
    mine.ij = 0xFFFFFFFFu;
    for (uint32_t idx = lane; idx < nt; idx += 32) {
        const Cand c = R.cand[idx];
        Best b;
        b.p = c.v;
        b.k = c.k & kOverlapMask;
        b.ij = c.ij;
        if (better(b, mine)) mine = b;
    }
    Best best = warp_best(mine);
    bool dense = R.overflow != 0u;
    bool ship = false;
    if (zero.ij != 0xFFFFFFFFu) {
        // a cell that is exactly 0.0 on both sides beats every non-zero p -- unless that p is itself a handful of
        // subnormal quanta, where host and device may disagree on which cells are zero at all
        if (R.pmin <= kTieAbs) dense = true;
        best = zero;
    } else if (best.ij == 0xFFFFFFFFu || !(best.p < 1.0)) {
        dense = true;  // no cell beats the cells the reference short-circuits to p = 1.0: needs the dense path
    } else if (nt > 1) {
        // more than one cell inside the ambiguity window: harmless iff all of them are the same (K, n, k)
        const uint32_t bK = R.s_c1[best.ij >> 16], bn = P.c2[best.ij & 0xFFFFu];
        bool differs = false;
        for (uint32_t idx = lane; idx < nt; idx += 32) {
            const Cand c = R.cand[idx];
            differs = differs || (c.k & kOverlapMask) != best.k || R.s_c1[c.ij >> 16] != bK || P.c2[c.ij & 0xFFFFu] != bn;
        }
        ship = __any_sync(kFull, differs);
    }
    if (ship) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&O.summary->n_tie_entries, nt);
        base = __shfl_sync(kFull, base, 0);
        if (base + nt > O.tie_cap) {
            // pool exhausted: the dense path settles this task; the part of the reservation that lies inside the pool is
            // marked void so the host does not read stale entries there
            dense = true;
            for (uint32_t idx = lane; idx < nt; idx += 32)
                if (base + idx < O.tie_cap) O.ties[base + idx].task = 0xFFFFFFFFu;
        } else {
            for (uint32_t idx = lane; idx < nt; idx += 32) {
                const Cand c = R.cand[idx];
                TieEntry t;
                t.task = (uint32_t)task;
                t.ij = c.ij;
                t.k = c.k & kOverlapMask;
                O.ties[base + idx] = t;
            }
        }
    }
    if (dense) {
        if (lane == 0) {
            O.full_list[atomicAdd(&O.summary->n_full, 1u)] = (uint32_t)task;
            atomicAdd(&O.summary->n_done, 1u);
        }
        __syncwarp();
        return;
    }
    if (lane == 0) {
        const uint32_t bi = best.ij >> 16, bj = best.ij & 0xFFFFu;
        dto_b200_record r;
        r.rank1 = P.thr1[bi];
        r.rank2 = P.thr2[bj];
        r.set1_len = R.s_c1[bi];
        r.set2_len = P.c2[bj];
        r.intersection_size = best.k;
        r.flags = record_flags;  // a shipped tie set is settled by the host, which rewrites the record (TIE_RESOLVED)
        r.population_size = P.N;
        r.pvalue = best.p;
        out[task] = r;
        atomicAdd(&O.summary->n_done, 1u);
    }
    // per-lane counters -> one atomic per warp
    uint32_t ne = R.n_eval, nr = R.n_refine;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ne += __shfl_xor_sync(kFull, ne, off);
        nr += __shfl_xor_sync(kFull, nr, off);
    }
    if (lane == 0) {
        atomicAdd(&counters[0], (unsigned long long)ne);
        atomicAdd(&counters[1], (unsigned long long)R.n_level2);
        atomicAdd(&counters[2], (unsigned long long)nr);
        if (task_stats) {
            uint32_t *ts = task_stats + (size_t)kTaskStatWords * task;
            ts[0] = R.n_level2;
            ts[1] = nr;
            ts[2] = ne;
            ts[3] = (uint32_t)((clock64() - t_begin) >> 4);
            ts[4] = (uint32_t)level;
            ts[5] = nt;
            ts[6] = ts[7] = 0;
        }
    }
    __syncwarp();
}

// State of one permutation's row loop that survives a trip through the rare path (kept in local memory between calls of
// scan_rows, in registers inside it).
template <int CH>
struct RowState {
    uint32_t qn;             // warp-uniform copy of the queue length (refreshed only after rows that pushed something)
    uint32_t g, pend;        // first list-1 position of the current 32-position group; lanes whose element is not binned yet
    int i, level;            // next row; screen level
};

__device__ __forceinline__ void ring_issue_chunk(const uint16_t *row, uint32_t n_chunks, uint32_t ring_addr, int lane,
                                                 uint32_t c) {
    if (c < n_chunks) {
        const uint16_t *src = row + (size_t)c * kChunk + lane * 4;
        const uint32_t dst = ring_addr + ((c & 3u) * kChunk + lane * 4) * 2;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// The row loop of one permutation, from row st.i until the screen queue holds a full chunk (st.qn >= 32) or the last
// row is done.  A function of its own ON PURPOSE: ptxas allocates registers per function, so the loop keeps its 2 x CH/2
// words of column state in registers no matter what the per-task prologue and the rare path of the kernel need.
template <int CH, bool SWAR>
__device__ __noinline__ void scan_rows(const Problem &P, RowState<CH> &st, const uint16_t *__restrict__ row,
                                       unsigned char *wbase, const uint32_t *s_c1) {
    using L = ScanLayout<CH>;
    constexpr int CHP = L::CHP;
    constexpr int NP = CH / 2;
    const int lane = threadIdx.x & 31;
    uint32_t *D = reinterpret_cast<uint32_t *>(wbase);
    uint32_t *Qij = reinterpret_cast<uint32_t *>(wbase + L::d_bytes);
    uint16_t *Qk = reinterpret_cast<uint16_t *>(Qij + L::QCAP);
    uint32_t *qcnt = reinterpret_cast<uint32_t *>(wbase + L::d_bytes + L::q_bytes);
    const uint16_t *ring = reinterpret_cast<const uint16_t *>(wbase + L::d_bytes + L::q_bytes + 16 + (size_t)L::CAP * sizeof(Cand));
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t n_chunks = P.pb_stride / kChunk;
    uint32_t qn = st.qn, g = st.g;
    int i = st.i;
    const int level = st.level;
    // The partner slots stream through the warp in groups of 32 consecutive list-1 positions, one element per lane,
    // independent of the row structure: a lane decodes its element ONCE (histogram word, which half) and bins it as soon
    // as the row loop reaches the row the position belongs to.  Rows shorter than a group -- most rows: the threshold
    // series grows by 1 % per step -- then cost one predicated atomic instead of a ring read, a decode and a loop each.
    // cp.async ring: chunk q (kChunk positions) is consumed while q+1 and q+2 are in flight or landed.
    auto enter_chunk = [&](uint32_t q) {
        ring_issue_chunk(row, n_chunks, ring_addr, lane, q + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
    };
    uint32_t cpos = g + (uint32_t)lane, coff, cadd;
    bool pend = ((st.pend >> lane) & 1u) != 0u;
    auto load_elem = [&]() {
        // kNoSlot (partner beyond the last list-2 threshold) lands in a sink word past the histogram
        const uint32_t slot = ring[cpos & (kRing - 1)];
        coff = min(slot & 0xFFFCu, L::dummy_off);
        cadd = (slot & 1u) * 0xFFFFu + 1u;
    };
    load_elem();
    // this lane's words of the current row of critical overlaps (level `level`, row i), advanced row by row
    const uint4 *__restrict__ kr =
        reinterpret_cast<const uint4 *>(P.kcrit + ((size_t)level * P.T1 + i) * P.T2pad) + lane;
    const uint32_t kr_step = (uint32_t)P.T2pad >> 3;  // 16-byte vectors per row
    constexpr int KV = kcrit_vecs(CH);
    const int T1 = P.T1;  // P lives behind a generic pointer here: read the loop bound once, not once per row
    uint4 *D4 = reinterpret_cast<uint4 *>(D + lane * CHP);  // this lane's column pairs, 4 per 128-bit vector
    {
        for (; i < T1; ++i) {
            const uint32_t hi = s_c1[i];
            // critical overlaps of this row at the current screen level: ceil(CH/8) coalesced 128-bit loads per lane, in
            // flight during the scatter.  (Requesting row i+1 one row ahead -- a twice-unrolled body with two register
            // sets -- was measured 13 % SLOWER at N = 20 000, like the software pipelining tried in round 1.)
            uint32_t kc2[KV * 4];
#pragma unroll
            for (int v = 0; v < KV; ++v) {
                const uint4 t = __ldg(kr + v * 32);
                kc2[4 * v] = t.x, kc2[4 * v + 1] = t.y, kc2[4 * v + 2] = t.z, kc2[4 * v + 3] = t.w;
            }
            kr += kr_step;
            // (1) bin this row's genes: every element of the stream below the row's end position, privatised per warp
            for (;;) {
                if (pend && cpos < hi) {
                    atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(D) + coff), cadd);
                    pend = false;
                }
                if (g + 32 > hi) break;  // the group reaches past this row: its remaining elements belong to later rows
                g += 32;
                cpos += 32;
                if ((g & (kChunk - 1)) == 0) enter_chunk(g / kChunk);
                load_elem();
                pend = true;
            }
            __syncwarp();
            // (2) 2-D inclusive prefix, everything packed (two 16-bit counts per word; no field ever exceeds 65534).  The
            // histogram is CUMULATIVE over the rows -- it is never cleared inside a permutation -- so word (x, y) of a column
            // pair already holds the column totals of all rows so far and the prefix over the columns is the overlap k(i, .)
            // itself: no per-row zeroing stores, no accumulators carried from row to row.  d = (x, y) of a column pair ->
            // d * 0x10001 = (x, x + y); the running total rides in both halves.  Two independent half-length chains, the
            // second one offset by the first one's total afterwards.
            constexpr int NV = (NP + 3) / 4, HA = NP / 2;
            uint32_t d[NV * 4];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const uint4 t = D4[v];
                d[4 * v] = t.x, d[4 * v + 1] = t.y, d[4 * v + 2] = t.z, d[4 * v + 3] = t.w;
            }
            uint32_t kcur2[NP];  // overlap of this lane's columns in this row, without the lanes to the left
            uint32_t runA = 0, runB = 0;
#pragma unroll
            for (int q = 0; q < HA; ++q) {
                const uint32_t t = d[q] * 0x10001u + runA;
                kcur2[q] = t;
                runA = __byte_perm(t, 0u, 0x3232);
            }
#pragma unroll
            for (int q = HA; q < NP; ++q) {
                const uint32_t t = d[q] * 0x10001u + runB;
                kcur2[q] = t + runA;
                runB = __byte_perm(t, 0u, 0x3232);
            }
            const uint32_t run2 = runA + runB;  // lane total in both halves
            uint32_t inc2 = run2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(kFull, inc2, o);
                if (lane >= o) inc2 += v;
            }
            const uint32_t koff2 = inc2 - run2;  // overlap contributed by the lanes to the left, in both halves
            // (3) screen: only k >= kcrit can have p <= tau_level.  Packed compare of both 16-bit fields at once, branch-free;
            // `ge2(q)` has bit 15 / bit 31 set iff the low / high column of pair q passes.
            //  * SWAR (every set size <= 32 766): field = k + 0x8000 - kcrit keeps its top bit iff k >= kcrit (all values
            //    < 0x8000, "never" = 0x7FFF, so no borrow crosses the fields): 2 instructions per pair;
            //  * longer lists (full 16-bit fields, "never" = 0xFFFF): compare the low 15 bits the same way, then settle the
            //    top bits: a >= b  <=>  (a15 & ~b15) | (~(a15 ^ b15) & low15(a) >= low15(b)): 5 instructions per pair.
            // One warp vote per row; 8 % of the rows have a passing cell, and only then each pair is revisited.
            const uint32_t bias = koff2 + 0x80008000u;
            auto ge2 = [&](int q) -> uint32_t {
                if constexpr (SWAR) {
                    return kcur2[q] + bias - kc2[q];
                } else {
                    const uint32_t a = kcur2[q] + koff2, bq = kc2[q];
                    const uint32_t d = (a | 0x80008000u) - (bq & 0x7FFF7FFFu);
                    return (a & ~bq) | (~(a ^ bq) & d);
                }
            };
            uint32_t hit = 0;
#pragma unroll
            for (int q = 0; q < NP; ++q) hit |= ge2(q);
            if (__any_sync(kFull, (hit & 0x80008000u) != 0u)) {  // some lane has a passing cell
                if (hit & 0x80008000u) {  // revisit the pairs, each guarded by its own test
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const uint32_t h = ge2(q) & 0x80008000u;
                        if (h) {
                            const uint32_t kk = kcur2[q] + koff2;  // no carry between the halves: every k < 65536
                            if (h & 0x8000u) {
                                const uint32_t slot = atomicAdd(qcnt, 1u);
                                Qij[slot] = ((uint32_t)i << 16) | (uint32_t)(lane * CH + 2 * q);
                                Qk[slot] = (uint16_t)(kk & 0xFFFFu);
                            }
                            if (h >> 31) {
                                const uint32_t slot = atomicAdd(qcnt, 1u);
                                Qij[slot] = ((uint32_t)i << 16) | (uint32_t)(lane * CH + 2 * q + 1);
                                Qk[slot] = (uint16_t)(kk >> 16);
                            }
                        }
                    }
                }
                __syncwarp();
                qn = *qcnt;
            }
            if (qn >= 32) {
                ++i;
                break;
            }
        }
    }
    st.qn = qn;
    st.g = g;
    st.pend = __ballot_sync(kFull, pend);
    st.i = i;
}

template <int CH, bool SWAR>
__global__ void __launch_bounds__(kScanThreads, (CH > 32) ? 1 : kScanCtasPerSm)
scan_kernel(const __grid_constant__ Problem P, const uint16_t *__restrict__ pb, int n_tasks, int n_plain,
            uint32_t record_flags, dto_b200_record *__restrict__ out, const ScanOut O,
            unsigned long long *__restrict__ counters, uint32_t *__restrict__ task_stats) {
    using L = ScanLayout<CH>;
    constexpr int CHP = L::CHP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *s_c1 = reinterpret_cast<uint32_t *>(smem_raw);  // [T1] shared by the CTA
    const size_t c1_bytes = ((size_t)P.T1 * 4 + 15) & ~(size_t)15;
    unsigned char *wbase = smem_raw + c1_bytes + (size_t)warp * L::per_warp;
    uint32_t *D = reinterpret_cast<uint32_t *>(wbase);
    uint32_t *Qij = reinterpret_cast<uint32_t *>(wbase + L::d_bytes);
    uint16_t *Qk = reinterpret_cast<uint16_t *>(Qij + L::QCAP);
    uint32_t *qcnt = reinterpret_cast<uint32_t *>(wbase + L::d_bytes + L::q_bytes);
    Cand *cand = reinterpret_cast<Cand *>(wbase + L::d_bytes + L::q_bytes + 16);
    uint16_t *ring = reinterpret_cast<uint16_t *>(wbase + L::d_bytes + L::q_bytes + 16 + (size_t)L::CAP * sizeof(Cand));
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring);

    for (int x = threadIdx.x; x < P.T1; x += blockDim.x) s_c1[x] = P.c1[x];
    __syncthreads();

    // dynamic scheduling: permutations differ in cost (how many cells pass the screen)
    for (;;) {
        int task = 0;
        if (lane == 0) task = (int)atomicAdd(reinterpret_cast<unsigned int *>(&counters[7]), 1u);
        task = __shfl_sync(kFull, task, 0);
        if (task >= n_tasks) break;
        const uint16_t *__restrict__ row = pb + (size_t)task * P.pb_stride;
        const long long t_begin = task_stats ? clock64() : 0;
        for (int x = lane; x < 32 * CHP; x += 32) D[x] = 0;
        if (lane == 0) *qcnt = 0;
        __syncwarp();

        Rare R;
        R.theta = CUDART_INF;
        R.pmin = CUDART_INF;
        R.overflow = 0u;
        R.zero.p = 0.0;
        R.zero.k = 0;
        R.zero.ij = 0xFFFFFFFFu;
        R.ncand = 0;
        R.n_level2 = R.n_eval = R.n_refine = 0;
        R.s_c1 = s_c1;
        R.Qij = Qij;
        R.Qk = Qk;
        R.qcnt = qcnt;
        R.cand = cand;
        R.lane = lane;

        RowState<CH> st;
        st.qn = st.g = 0;
        st.pend = kFull;
        st.i = st.level = 0;
        // partner-slot row staged through shared memory with cp.async: chunk c = kChunk positions, one 8 B copy per
        // lane; chunk 0 must have landed before the row loop starts, chunks 1 and 2 stay in flight (ring of 4 chunks)
        const uint32_t n_chunks = P.pb_stride / kChunk;
        ring_issue_chunk(row, n_chunks, ring_addr, lane, 0);
        ring_issue_chunk(row, n_chunks, ring_addr, lane, 1);
        ring_issue_chunk(row, n_chunks, ring_addr, lane, 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        // The row loop lives in scan_rows and is call-free: when the queue holds a full chunk it returns to the
        // (out-of-line) drain and is re-entered, so the column state is only saved/restored around that rare trip.
        while (st.i < P.T1) {
            scan_rows<CH, SWAR>(P, st, row, wbase, s_c1);
            if (st.qn >= 32) {
                st.level = drain_queue(P, R, false, st.level);
                __syncwarp();
                st.qn = *qcnt;
            }
        }
        const int level = st.level;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // the first n_plain tasks of a launch are unpermuted ones riding along (batched list pairs)
        finish_task(P, R, task, level, task < n_plain ? (record_flags & ~DTO_B200_FLAG_PERMUTED) : record_flags, out, O,
                    counters, task_stats, t_begin);
    }
}

// =====================================================================================================
// K3: dense grid
// =====================================================================================================
__global__ void full_hist_kernel(const Problem P, const uint16_t *__restrict__ pbrow, uint32_t *__restrict__ H) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n1_eff) return;
    const uint32_t slot = pbrow[j];
    const uint32_t b1 = P.bin1[j];
    if (slot == kNoSlot || b1 == kNoSlot) return;
    const uint32_t col = hist_slot_column(slot, P.CH);
    atomicAdd(&H[(size_t)b1 * P.T2 + col], 1u);
}

__global__ void full_prefix_cols_kernel(int T1, int T2, uint32_t *__restrict__ H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= T2) return;
    uint32_t run = 0;
    for (int i = 0; i < T1; ++i) {
        run += H[(size_t)i * T2 + j];
        H[(size_t)i * T2 + j] = run;
    }
}

__global__ void full_prefix_rows_kernel(int T1, int T2, uint32_t *__restrict__ H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T1) return;
    uint32_t run = 0;
    for (int j = 0; j < T2; ++j) {
        run += H[(size_t)i * T2 + j];
        H[(size_t)i * T2 + j] = run;
    }
}

__global__ void full_eval_kernel(const Problem P, const uint32_t *__restrict__ H, double *__restrict__ pv,
                                 double *__restrict__ logp) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= P.T1 * P.T2) return;
    const int i = cell / P.T2, j = cell % P.T2;
    const uint64_t K = P.c1[i], n = P.c2[j], k = H[cell];
    if (pv) pv[cell] = hypergeom_pvalue_exact(P.lf, P.N, K, n, k);
    if (logp) logp[cell] = hypergeom_log_pvalue(P.lf, P.N, K, n, k);
}

__global__ void __launch_bounds__(1024) full_argmin_kernel(const Problem P, const uint32_t *__restrict__ H,
                                                           const double *__restrict__ pv, uint32_t record_flags,
                                                           dto_b200_record *__restrict__ out,
                                                           uint32_t *__restrict__ best_cell) {
    __shared__ Best sb[32];
    Best best;
    best.p = CUDART_INF;
    best.k = 0;
    best.ij = 0xFFFFFFFFu;
    const int cells = P.T1 * P.T2;
    for (int cell = threadIdx.x; cell < cells; cell += blockDim.x) {
        Best b;
        b.p = pv[cell];
        b.k = H[cell];
        b.ij = ((uint32_t)(cell / P.T2) << 16) | (uint32_t)(cell % P.T2);
        if (better(b, best)) best = b;
    }
    best = warp_best(best);
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        Best b = sb[threadIdx.x];
        if (threadIdx.x >= (blockDim.x >> 5)) b.ij = 0xFFFFFFFFu;
        b = warp_best(b);
        if (threadIdx.x == 0) {
            const uint32_t bi = b.ij >> 16, bj = b.ij & 0xFFFFu;
            dto_b200_record r;
            r.rank1 = P.thr1[bi];
            r.rank2 = P.thr2[bj];
            r.set1_len = P.c1[bi];
            r.set2_len = P.c2[bj];
            r.intersection_size = b.k;
            r.flags = record_flags | DTO_B200_FLAG_PATH_FULL;
            r.population_size = P.N;
            r.pvalue = b.p;
            *out = r;
            if (best_cell) *best_cell = bi * (uint32_t)P.T2 + bj;
        }
    }
}

// Dense-path tie set: every cell whose device p lies inside the ambiguity window of the device minimum (plus the argmin
// itself), as {cell, overlap}; the host re-evaluates them with its libm and applies the reference tie-break.  When the
// minimum is exactly 0.0 the (possibly thousands of) other exact zeros are NOT listed: exp() underflows identically on
// both sides, so among them the integer tie-break of full_argmin_kernel is already the reference's; only cells a few
// subnormal quanta above zero are ambiguous then.
// count[0] = listed cells; when the minimum is exactly 0.0: count[1] = cells on the zero plateau, count[2] = those of them
// with the winning overlap (what the reference's two tie notices, optimize_main.rs:87-107, are about).
__global__ void full_collect_kernel(const Problem P, const uint32_t *__restrict__ H, const double *__restrict__ pv,
                                    const uint32_t *__restrict__ best_cell, uint32_t *__restrict__ count,
                                    uint2 *__restrict__ cells_out) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= P.T1 * P.T2) return;
    const uint32_t bc = *best_cell;
    const double pb = pv[bc], p = pv[cell];
    bool take = (uint32_t)cell == bc;
    if (!take) take = pb > 0.0 ? (p <= pb * (1.0 + kTieRel) + kTieAbs) : (p > 0.0 && p <= kTieAbs);
    if (take) cells_out[atomicAdd(count, 1u)] = make_uint2((uint32_t)cell, H[cell]);
    if (pb == 0.0 && p == 0.0) {
        atomicAdd(count + 1, 1u);
        if (H[cell] == H[bc]) atomicAdd(count + 2, 1u);
    }
}

// records[idx[x]] = patch[x]: host-resolved records written back into the device-resident record array
__global__ void patch_records_kernel(const uint32_t *__restrict__ idx, const dto_b200_record *__restrict__ patch, int n,
                                     dto_b200_record *__restrict__ records) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n) records[idx[x]] = patch[x];
}

__global__ void pvalues_kernel(const double *__restrict__ lf, const uint64_t *__restrict__ N,
                               const uint64_t *__restrict__ K, const uint64_t *__restrict__ n,
                               const uint64_t *__restrict__ k, size_t count, double *__restrict__ out) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= count) return;
    out[x] = hypergeom_pvalue_exact(lf, N[x], K[x], n[x], k[x]);
}

// =====================================================================================================
// roofline probes
// =====================================================================================================
__global__ void fp64_probe_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
    double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void hbm_copy_probe_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x)
        dst[x] = src[x];
}

// =====================================================================================================
// launchers
// =====================================================================================================
// opts a kernel into the device's maximum dynamic shared memory, once per device (bit d of `done` = device d is set up)
static cudaError_t set_max_smem_once(const void *kern, std::atomic<uint64_t> &done) {
    int dev = 0, v = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

template <int CH, bool SWAR>
static cudaError_t launch_scan_t(const Problem &P, const uint16_t *pb, int n_tasks, int n_plain, uint32_t flags,
                                 dto_b200_record *out, const ScanOut &status, unsigned long long *counters,
                                 uint32_t *task_stats, int grid, int warps, cudaStream_t st) {
    using L = ScanLayout<CH>;
    const size_t smem = (((size_t)P.T1 * 4 + 15) & ~(size_t)15) + (size_t)warps * L::per_warp;
    auto kern = scan_kernel<CH, SWAR>;
    // always the device maximum: the attribute is per function and device, and several contexts (host threads) may launch
    // the same instantiation with different sizes concurrently.  Set once per (instantiation, device).
    static std::atomic<uint64_t> done{0};
    cudaError_t e = set_max_smem_once(reinterpret_cast<const void *>(kern), done);
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, st>>>(P, pb, n_tasks, n_plain, flags, out, status, counters, task_stats);
    return cudaGetLastError();
}

size_t scan_smem_bytes(int CH, int T1, int warps) {
    const size_t chp = (size_t)hist_words(CH);
    const size_t d = 32 * chp * 4 + 16;
    const size_t q = ((size_t)(32 * CH + 32) * 6 + 15) & ~(size_t)15;
    const size_t per = d + q + 16 + (size_t)kCandCap * sizeof(Cand) + (size_t)kRing * 2;
    return (((size_t)T1 * 4 + 15) & ~(size_t)15) + (size_t)warps * per;
}

int pick_ch(int T2) {
    static const int opts[] = {2, 4, 8, 12, 16, 20, 24, 28, 32, 48, 64};
    const int need = (T2 + 31) / 32;
    for (int o : opts)
        if (o >= need) return o;
    return -1;
}

cudaError_t launch_scan(const Problem &P, const uint16_t *pb, int n_tasks, int n_plain, uint32_t flags, dto_b200_record *out,
                        const ScanOut &status, unsigned long long *counters, uint32_t *task_stats, int grid, int warps,
                        cudaStream_t st) {
#define DTO_CASE(X)                                                                                                 \
    case X:                                                                                                         \
        return P.never == 0x7FFFu                                                                                   \
                   ? launch_scan_t<X, true>(P, pb, n_tasks, n_plain, flags, out, status, counters, task_stats, grid, warps, st) \
                   : launch_scan_t<X, false>(P, pb, n_tasks, n_plain, flags, out, status, counters, task_stats, grid, warps, st);
    switch (P.CH) {
        DTO_CASE(2) DTO_CASE(4) DTO_CASE(8) DTO_CASE(12) DTO_CASE(16) DTO_CASE(20) DTO_CASE(24)
        DTO_CASE(28) DTO_CASE(32) DTO_CASE(48) DTO_CASE(64)
        default:
            return cudaErrorInvalidValue;
    }
#undef DTO_CASE
}

cudaError_t launch_build_kcrit(const Problem &P, uint16_t *kcrit, uint32_t *counts, uint2 *meta, cudaStream_t st) {
    const size_t words = (size_t)(P.levels + 1) * P.T1 * P.T2pad;
    fill_u16_kernel<<<1024, 256, 0, st>>>(kcrit, words, (uint16_t)P.never);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int cells = P.T1 * P.T2;
    e = cudaMemsetAsync(meta, 0, (size_t)cells * sizeof(uint2), st);
    if (e != cudaSuccess) return e;
    build_kcrit_kernel<<<(cells + 255) / 256, 256, 0, st>>>(P, kcrit, 0, counts, meta, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_scan_counts(uint32_t *counts, int n, unsigned long long *total_out, cudaStream_t st) {
    exclusive_scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n, total_out);
    return cudaGetLastError();
}

cudaError_t launch_fill_lptab(const Problem &P, const uint32_t *offsets, uint2 *meta, double *lptab, cudaStream_t st) {
    const int cells = P.T1 * P.T2;
    set_meta_offsets_kernel<<<(cells + 255) / 256, 256, 0, st>>>(cells, offsets, meta);
    build_kcrit_kernel<<<(cells + 255) / 256, 256, 0, st>>>(P, nullptr, 1, nullptr, meta, lptab);
    return cudaGetLastError();
}

// shared-memory bytes of the pairing kernel without the boundary list; ba_in_smem = false moves the n x 4 B
// (key16 | arrival) array to the per-CTA global scratch (long lists)
size_t sigma_smem_bytes(const Problem &P, int B1, int B2, bool ba_in_smem) {
    const uint32_t nmax = P.n1 > P.n2 ? P.n1 : P.n2;
    const uint32_t nmax8 = (nmax + 7) & ~7u;
    const int Bmax = B1 > B2 ? B1 : B2;
    size_t b = (((size_t)(1u << Bmax) + 1 + kScanTmp) * 4 + 15) & ~(size_t)15;
    b += (size_t)(P.pb_stride > nmax8 ? P.pb_stride : nmax8) * 2;
    b += (((size_t)P.T1 * 4) + 15) & ~(size_t)15;
    const bool identical = (P.n_common == P.n1 && P.n_common == P.n2);
    if (!identical) b += (((size_t)P.n_common * 2) + 15) & ~(size_t)15;
    if (ba_in_smem) b += (size_t)nmax8 * 4;
    return (b + 15) & ~(size_t)15;
}

// u32 words of global scratch one CTA of the pairing kernel needs
size_t sigma_scratch_words(const Problem &P) {
    const uint32_t nmax = P.n1 > P.n2 ? P.n1 : P.n2;
    return (size_t)4 * ((nmax + 7) & ~7u);
}

int pick_bucket_bits(uint32_t n) {
    int lg = 0;
    while ((1ull << lg) < n) ++lg;
    int B = lg - 1;  // 1..2 elements per bucket
    if (B < 1) B = 1;
    if (B > 14) B = 14;
    return B;
}

// ctas_per_sm: 0 = choose (two 512-thread CTAs per SM on the row-wise path when they fit, else one 1024-thread CTA),
// 1 / 2 = force.  *grid_unit_out = CTAs one SM holds (the caller sizes the grid in multiples of sm_count x that).
cudaError_t launch_sigma_sort(const Problem &P, uint64_t seed, const uint64_t *seeds, uint32_t seg, uint64_t first_id,
                              int n_tasks, uint16_t *pb, uint32_t *pairing_out, uint32_t *scratch, size_t smem_limit,
                              int sm_count, int ctas_per_sm, cudaStream_t st) {
    const int B1 = pick_bucket_bits(P.n1), B2 = pick_bucket_bits(P.n2);
    if (!scratch) return cudaErrorInvalidValue;
    const bool rowwise = (P.n_common == P.n1 && P.n_common == P.n2) && pairing_out == nullptr;
    const uint32_t nmax = P.n1 > P.n2 ? P.n1 : P.n2;
    // two CTAs per SM: each may use half of the SM's shared memory minus the 1 KB the system reserves per CTA
    const size_t half_limit = ((size_t)228 * 1024) / 2 - 1024;
    // Measured on a B200 (profiles/r02_sigma_ctas_ab.json): two CTAs per SM win 24 % at N = 6 000, where the (key16 |
    // arrival) words of both fit in shared memory; at N = 20 000 they only fit with those words in the L2 scratch, and that
    // costs more (+7 %) than the overlap of the two CTAs' barriers gains -- so: two CTAs only with everything on chip.
    int per_sm = 1;
    const bool fits_two_lean = sigma_smem_bytes(P, B1, B2, false) + 1024 <= half_limit && (1u << B2) <= 32u * 512u;
    const bool fits_two_onchip = sigma_smem_bytes(P, B1, B2, true) + 1024 <= half_limit && (1u << B2) <= 32u * 512u;
    if (rowwise && ((ctas_per_sm == 0 && fits_two_onchip) || (ctas_per_sm == 2 && fits_two_lean))) per_sm = 2;
    if (ctas_per_sm == 2 && per_sm != 2) return cudaErrorInvalidValue;
    const size_t limit = per_sm == 2 ? half_limit : smem_limit;
    const bool ba_in_smem = sigma_smem_bytes(P, B1, B2, true) <= limit;
    const size_t base = sigma_smem_bytes(P, B1, B2, ba_in_smem);
    if (base > limit) return cudaErrorInvalidValue;
    // whatever shared memory is left holds the boundary-bucket list (it spills to the global scratch beyond that)
    const size_t list_cap = std::min<size_t>((limit - base) / 4 & ~(size_t)3, (nmax + 7) & ~7u);
    const size_t smem = base + list_cap * 4;
    const int grid = std::min(n_tasks, sm_count * 4);  // a multiple of sm_count x per_sm; the scratch holds sm_count x 4 CTAs
    static std::atomic<uint64_t> done[4];
    cudaError_t e;
#define DTO_SIGMA(SCR, NTH, SLOT)                                                                                         \
    {                                                                                                                      \
        auto kern = sigma_sort_kernel<SCR, NTH>;                                                                           \
        e = set_max_smem_once(reinterpret_cast<const void *>(kern), done[SLOT]);                                           \
        if (e != cudaSuccess) return e;                                                                                    \
        kern<<<grid, NTH, smem, st>>>(P, seed, seeds, seg ? seg : 1u, first_id, n_tasks, B1, B2, pb, pairing_out, scratch, \
                                      (uint32_t)list_cap);                                                                 \
    }
    if (per_sm == 2) {
        if (ba_in_smem) DTO_SIGMA(false, 512, 0) else DTO_SIGMA(true, 512, 1)
    } else {
        if (ba_in_smem) DTO_SIGMA(false, 1024, 2) else DTO_SIGMA(true, 1024, 3)
    }
#undef DTO_SIGMA
    return cudaGetLastError();
}

cudaError_t launch_compose(const Problem &P, const uint32_t *perm1, const uint32_t *perm2, const int32_t *slot2_maps,
                           int n_tasks, uint32_t *inv_scratch, int *err_flag, uint16_t *pb, cudaStream_t st) {
    const int th = 256;
    if (perm1 || perm2) {  // an inverse slot no entry writes (duplicates elsewhere in the row) stays detectably invalid
        const uint32_t nmax = P.n1 > P.n2 ? P.n1 : P.n2;
        cudaError_t e = cudaMemsetAsync(inv_scratch, 0xFF, (size_t)nmax * n_tasks * 4, st);
        if (e != cudaSuccess) return e;
    }
    if (perm1) {  // validate perm1 is a permutation (scratch reused)
        const size_t tot = (size_t)P.n1 * n_tasks;
        invert_perm_kernel<<<(unsigned)((tot + th - 1) / th), th, 0, st>>>(perm1, P.n1, n_tasks, inv_scratch, err_flag);
        check_perm_kernel<<<(unsigned)((tot + th - 1) / th), th, 0, st>>>(perm1, P.n1, n_tasks, inv_scratch, err_flag);
    }
    if (perm2) {
        const size_t tot = (size_t)P.n2 * n_tasks;
        invert_perm_kernel<<<(unsigned)((tot + th - 1) / th), th, 0, st>>>(perm2, P.n2, n_tasks, inv_scratch, err_flag);
        check_perm_kernel<<<(unsigned)((tot + th - 1) / th), th, 0, st>>>(perm2, P.n2, n_tasks, inv_scratch, err_flag);
    }
    const size_t tot = (size_t)P.pb_stride * n_tasks;
    compose_pairing_kernel<<<(unsigned)((tot + th - 1) / th), th, 0, st>>>(P, perm1, perm2 ? inv_scratch : nullptr,
                                                                         slot2_maps, n_tasks, pb);
    return cudaGetLastError();
}

cudaError_t launch_full_grid(const Problem &P, const uint16_t *pbrow, uint32_t *H, double *pv, double *logp,
                             cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(H, 0, (size_t)P.T1 * P.T2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (P.n1_eff) full_hist_kernel<<<(P.n1_eff + 255) / 256, 256, 0, st>>>(P, pbrow, H);
    full_prefix_cols_kernel<<<(P.T2 + 127) / 128, 128, 0, st>>>(P.T1, P.T2, H);
    full_prefix_rows_kernel<<<(P.T1 + 127) / 128, 128, 0, st>>>(P.T1, P.T2, H);
    if (pv || logp) full_eval_kernel<<<(P.T1 * P.T2 + 127) / 128, 128, 0, st>>>(P, H, pv, logp);
    return cudaGetLastError();
}

cudaError_t launch_full_argmin(const Problem &P, const uint32_t *H, const double *pv, uint32_t flags,
                               dto_b200_record *out, uint32_t *best_cell, cudaStream_t st) {
    full_argmin_kernel<<<1, 1024, 0, st>>>(P, H, pv, flags, out, best_cell);
    return cudaGetLastError();
}

cudaError_t launch_full_collect(const Problem &P, const uint32_t *H, const double *pv, const uint32_t *best_cell,
                                uint32_t *count, uint2 *cells_out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(count, 0, 12, st);
    if (e != cudaSuccess) return e;
    full_collect_kernel<<<(P.T1 * P.T2 + 255) / 256, 256, 0, st>>>(P, H, pv, best_cell, count, cells_out);
    return cudaGetLastError();
}

cudaError_t launch_patch_records(const uint32_t *idx, const dto_b200_record *patch, int n, dto_b200_record *records,
                                 cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    patch_records_kernel<<<(n + 127) / 128, 128, 0, st>>>(idx, patch, n, records);
    return cudaGetLastError();
}

cudaError_t launch_pvalues(const double *lf, const uint64_t *N, const uint64_t *K, const uint64_t *n,
                           const uint64_t *k, size_t count, double *out, cudaStream_t st) {
    pvalues_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(lf, N, K, n, k, count, out);
    return cudaGetLastError();
}

cudaError_t launch_fp64_probe(double *out, int blocks, int threads, int iters, cudaStream_t st) {
    fp64_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
    return cudaGetLastError();
}

cudaError_t launch_hbm_probe(const void *src, void *dst, size_t bytes, int blocks, cudaStream_t st) {
    hbm_copy_probe_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst),
                                                  bytes / 16);
    return cudaGetLastError();
}

}  // namespace dto
