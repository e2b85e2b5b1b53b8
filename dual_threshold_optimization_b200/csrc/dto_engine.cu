// dto_engine.cu -- low-level C ABI (integer ids): context, problem upload, batched task execution.
// Replaces the per-task body of run::run_single_node's worker loop (src/run/single_node.rs:114-123, i.e.
// dto::optimize -> process_threshold_pairs) with batched kernel launches.  No CPU fallback exists: every
// entry point fails with DTO_B200_ERR_CUDA when the device is unusable.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include "dto_host_math.hpp"
#include "dto_internal.hpp"
#include "dto_kernels.cuh"

namespace dto {

thread_local std::string g_last_error;

// process-wide totals over every context, pooled ones included (dto_b200_process_totals): what a caller of the host layer
// (run_tasks / run_pairs own their contexts) can still observe about launches and copies
std::atomic<uint64_t> g_total_launches{0}, g_total_h2d{0}, g_total_d2h{0};

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(DTO_B200_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                        __LINE__, #expr);                                                                   \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace dto

using namespace dto;

struct dto_b200_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool has_problem = false;
    Problem P{};
    // problem tables
    DevBuf d_c1, d_c2, d_thr1, d_thr2, d_lf, d_rowA, d_colB, d_kcrit, d_dslot2, d_bin1, d_bin2, d_slot2, d_meta, d_lptab, d_counts;
    // batch state
    DevBuf d_pb, d_records, d_counters, d_H, d_pv, d_logp, d_perm1, d_perm2, d_inv, d_err, d_pair, d_minp, d_tstats,
        d_words;
    // what a scan launch reports besides records: summary word block, ids of tasks for the dense path, unsettled tie sets
    DevBuf d_summary, d_full_list, d_ties, d_patch_idx, d_patch_rec, d_best_cell, d_collect;
    DevBuf d_slotmaps, d_seeds;  // batched list pairs: per-pair gene maps (unpermuted tasks) and per-pair Philox seeds
    PinnedBuf h_summary, h_ties, h_patch;
    std::vector<uint32_t> h_c1, h_c2, h_thr1, h_thr2;  // host copies of the current problem (tie resolution builds records)
    std::vector<uint32_t> h_ranks1, h_ranks2;          // ... and its ranks (do two list pairs share one rank structure?)
    bool opt_task_stats = false;
    bool opt_swar = true;
    bool opt_table_cache = true;
    std::vector<double> lf_host;  // ln_factorial(0..lf_N): a pure function of the population, kept across problems
    uint64_t lf_N = ~0ull;
    // The screen and log-p tables are a pure function of (population, set sizes per threshold, levels, packing): list
    // pairs that share them (e.g. tie-free ranks 1..n of equal length, config 4) reuse the tables of the previous problem.
    bool tab_valid = false;
    uint64_t tab_N = 0, tab_entries = 0;
    int tab_levels = 0;
    uint32_t tab_never = 0;
    std::vector<uint32_t> tab_c1, tab_c2;
    int last_batch_n = 0;
    PinnedBuf h_records, h_stage;
    // options
    int opt_batch = 0;  // 0 = auto
    int opt_warps = kScanThreads / 32;
    int opt_levels = 32;
    int opt_sigma_ctas = 0;  // pairing kernel: 0 = choose, 1 = one 1024-thread CTA per SM, 2 = two 512-thread CTAs per SM
    dto_b200_stats stats{};
};

namespace {

inline void count_launches(dto_b200_ctx *ctx, uint64_t n) {
    ctx->stats.kernel_launches += n;
    g_total_launches.fetch_add(n, std::memory_order_relaxed);
}
inline void count_h2d(dto_b200_ctx *ctx, uint64_t bytes) {
    ctx->stats.h2d_bytes += bytes;
    g_total_h2d.fetch_add(bytes, std::memory_order_relaxed);
}
inline void count_d2h(dto_b200_ctx *ctx, uint64_t bytes) {
    ctx->stats.d2h_bytes += bytes;
    g_total_d2h.fetch_add(bytes, std::memory_order_relaxed);
}

int bind(dto_b200_ctx *ctx) {
    if (!ctx) return fail(DTO_B200_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return DTO_B200_OK;
}

int auto_batch(const dto_b200_ctx *ctx) {
    if (ctx->opt_batch > 0) return ctx->opt_batch;
    // One warp works through whole permutations, so a launch ends with a ragged last wave of about half a permutation's
    // run time: 32 permutations per resident warp keep that tail under 2 % (8 per warp measured 8 % slower at N = 20 000),
    // bounded by ~6 GiB of partner-slot rows.
    const size_t row_bytes = (size_t)ctx->P.pb_stride * 2;
    size_t b = (size_t)ctx->sm_count * 2 * ctx->opt_warps * 32;
    const size_t cap = ((size_t)6 << 30) / (row_bytes ? row_bytes : 1);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// The reference's own evaluation of one cell on this host (statrs order, host libm): see dto_host_math.hpp.
double host_p(const dto_b200_ctx *ctx, uint32_t K, uint32_t n, uint32_t k) {
    return host_hypergeom_pvalue_exact(ctx->lf_host.data(), ctx->P.N, K, n, k);
}

struct HostCell {
    uint32_t ij, k;
};

// optimize_main.rs:73-116 over a handful of cells, with p re-evaluated on the host: smallest p (exact ==), then largest
// overlap, then smallest (rank1, rank2) == smallest (row, column).  Cells with equal (K, n, k) are evaluated once.
dto_b200_record resolve_on_host(dto_b200_ctx *ctx, const HostCell *cells, size_t n_cells, uint32_t flags,
                                uint32_t *n_min_out = nullptr, uint32_t *n_maxk_out = nullptr) {
    const int T2 = ctx->P.T2;
    (void)T2;
    double best_p = INFINITY;
    uint32_t best_k = 0, best_ij = 0xFFFFFFFFu;
    uint32_t lastK = ~0u, lastn = ~0u, lastk = ~0u;
    double lastp = 0.0;
    uint32_t n_min = 0, n_maxk = 0;  // cells sharing the minimum p exactly / of those, sharing the largest overlap
    for (size_t x = 0; x < n_cells; ++x) {
        const uint32_t i = cells[x].ij >> 16, j = cells[x].ij & 0xFFFFu, k = cells[x].k;
        const uint32_t K = ctx->h_c1[i], n = ctx->h_c2[j];
        double p;
        if (K == lastK && n == lastn && k == lastk) {
            p = lastp;
        } else {
            p = host_p(ctx, K, n, k);
            lastK = K, lastn = n, lastk = k, lastp = p;
        }
        if (best_ij == 0xFFFFFFFFu || p < best_p) n_min = 1, n_maxk = 1;
        else if (p == best_p) {
            ++n_min;
            if (k > best_k) n_maxk = 1;
            else if (k == best_k) ++n_maxk;
        }
        const bool better = best_ij == 0xFFFFFFFFu || p < best_p || (p == best_p && (k > best_k || (k == best_k && cells[x].ij < best_ij)));
        if (better) best_p = p, best_k = k, best_ij = cells[x].ij;
    }
    if (n_min_out) *n_min_out = n_min;
    if (n_maxk_out) *n_maxk_out = n_maxk;
    ctx->stats.tie_cells_host += n_cells;
    const uint32_t bi = best_ij >> 16, bj = best_ij & 0xFFFFu;
    dto_b200_record r;
    r.rank1 = ctx->h_thr1[bi];
    r.rank2 = ctx->h_thr2[bj];
    r.set1_len = ctx->h_c1[bi];
    r.set2_len = ctx->h_c2[bj];
    r.intersection_size = best_k;
    r.flags = flags | DTO_B200_FLAG_HOST_PVALUE;
    r.population_size = ctx->P.N;
    r.pvalue = best_p;
    return r;
}

// Dense path for ONE task whose partner-slot row is pbrow: every cell evaluated on the device in statrs order, device
// argmin, then the cells inside the ambiguity window of that minimum go to the host, which settles the optimum exactly
// as the reference would (host libm).  The record (host-evaluated p) is returned and, if d_dst != nullptr, also stored there.
int dense_task(dto_b200_ctx *ctx, const uint16_t *pbrow, uint32_t flags, dto_b200_record *d_dst, dto_b200_record *h_out) {
    const Problem &P = ctx->P;
    const size_t cells = (size_t)P.T1 * P.T2;
    CUDA_TRY(ctx->d_H.ensure(cells * 4));
    CUDA_TRY(ctx->d_pv.ensure(cells * 8));
    CUDA_TRY(ctx->d_best_cell.ensure(16 + sizeof(dto_b200_record)));
    CUDA_TRY(ctx->d_collect.ensure(cells * sizeof(uint2)));
    uint32_t *d_bc = ctx->d_best_cell.as<uint32_t>();  // [0] argmin cell, [1] collect count, then a scratch record
    dto_b200_record *d_rec = reinterpret_cast<dto_b200_record *>(d_bc + 4);
    CUDA_TRY(launch_full_grid(P, pbrow, ctx->d_H.as<uint32_t>(), ctx->d_pv.as<double>(), nullptr, ctx->stream));
    CUDA_TRY(launch_full_argmin(P, ctx->d_H.as<uint32_t>(), ctx->d_pv.as<double>(), flags, d_rec, d_bc, ctx->stream));
    CUDA_TRY(launch_full_collect(P, ctx->d_H.as<uint32_t>(), ctx->d_pv.as<double>(), d_bc, d_bc + 1, ctx->d_collect.as<uint2>(), ctx->stream));
    count_launches(ctx, 6);
    ctx->stats.tasks_full += 1;
    uint32_t head[4] = {0, 0, 0, 0};  // argmin cell, listed cells, zero-plateau cells, of those with the winning overlap
    CUDA_TRY(cudaMemcpyAsync(head, d_bc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint32_t cnt = head[1];
    if (cnt == 0 || cnt > cells) return fail(DTO_B200_ERR_CUDA, "dense path returned an invalid tie count (%u)", cnt);
    std::vector<uint2> got(cnt);
    CUDA_TRY(cudaMemcpy(got.data(), ctx->d_collect.p, (size_t)cnt * sizeof(uint2), cudaMemcpyDeviceToHost));
    count_d2h(ctx, 8 + (uint64_t)cnt * sizeof(uint2));
    std::vector<HostCell> hc(cnt);
    for (uint32_t x = 0; x < cnt; ++x) {
        hc[x].ij = ((got[x].x / (uint32_t)P.T2) << 16) | (got[x].x % (uint32_t)P.T2);
        hc[x].k = got[x].y;
    }
    // atomics fill the list in arbitrary order; equal (K, n, k) runs are evaluated once when adjacent
    std::sort(hc.begin(), hc.end(), [&](const HostCell &a, const HostCell &b) {
        const uint32_t Ka = ctx->h_c1[a.ij >> 16], Kb = ctx->h_c1[b.ij >> 16];
        if (Ka != Kb) return Ka < Kb;
        const uint32_t na = ctx->h_c2[a.ij & 0xFFFFu], nb = ctx->h_c2[b.ij & 0xFFFFu];
        if (na != nb) return na < nb;
        if (a.k != b.k) return a.k < b.k;
        return a.ij < b.ij;
    });
    uint32_t n_min = 0, n_maxk = 0;
    dto_b200_record r = resolve_on_host(ctx, hc.data(), hc.size(), flags | DTO_B200_FLAG_PATH_FULL | (cnt > 1 ? DTO_B200_FLAG_TIE_RESOLVED : 0u),
                                        &n_min, &n_maxk);
    if (cnt > 1) ctx->stats.tasks_tie_resolved += 1;
    if (r.pvalue == 0.0 && head[2] > 0) {
        // zero plateau: the listed cells are the argmin and the few-quanta neighbours; the plateau itself was counted on
        // the device (exp() underflows identically on both sides)
        n_min = head[2] + (n_min > 0 ? n_min - 1 : 0);
        n_maxk = std::max(n_maxk, head[3]);
    }
    // the situations in which the reference prints its two notices (optimize_main.rs:85-107)
    if (n_min > 1) r.flags |= DTO_B200_FLAG_TIE_MINP;
    if (n_min > 1 && n_maxk > 1) r.flags |= DTO_B200_FLAG_TIE_OVERLAP;
    if (d_dst) {
        CUDA_TRY(cudaMemcpy(d_dst, &r, sizeof(r), cudaMemcpyHostToDevice));
        count_h2d(ctx, sizeof(r));
    }
    if (h_out) *h_out = r;
    return DTO_B200_OK;
}

// Runs the scan over n tasks whose partner-slot rows are already in ctx->d_pb; results land in ctx->d_records[0..n).
// Tasks whose tie set needs the host libm (optimize_main.rs:73-80 compares p with ==) are settled here and patched
// into d_records; degenerate tasks (minimum p >= 1, tie set larger than the buffer) go through the dense path.
// flags_unperm_first: the first `n_unperm_first` tasks carry no PERMUTED flag (batched list pairs).
int run_tasks(dto_b200_ctx *ctx, int n, uint32_t flags, int n_plain = 0) {
    const Problem &P = ctx->P;
    CUDA_TRY(ctx->d_records.ensure((size_t)n * sizeof(dto_b200_record)));
    const uint32_t tie_cap = (uint32_t)std::max<size_t>(4096, (size_t)n / 2);
    CUDA_TRY(ctx->d_summary.ensure(sizeof(ScanSummary)));
    CUDA_TRY(ctx->h_summary.ensure(sizeof(ScanSummary)));
    CUDA_TRY(ctx->d_full_list.ensure((size_t)n * 4));
    CUDA_TRY(ctx->d_ties.ensure((size_t)tie_cap * sizeof(TieEntry)));
    CUDA_TRY(cudaMemsetAsync(ctx->d_summary.p, 0, sizeof(ScanSummary), ctx->stream));
    int warps = ctx->opt_warps;
    const int ctas_per_sm = (P.CH > 32) ? 1 : kScanCtasPerSm;
    while (warps > 1 && scan_smem_bytes(P.CH, P.T1, warps) + 1024 > (size_t)(228 * 1024) / ctas_per_sm) --warps;
    if (scan_smem_bytes(P.CH, P.T1, warps) > ctx->smem_optin)
        return fail(DTO_B200_ERR_UNSUPPORTED, "scan kernel does not fit shared memory (T1=%d T2=%d)", P.T1, P.T2);
    const int ctas_needed = (n + warps - 1) / warps;
    const int grid = std::min(ctas_needed, ctx->sm_count * ctas_per_sm);
    if (ctx->opt_task_stats) {
        CUDA_TRY(ctx->d_tstats.ensure((size_t)n * 4 * kTaskStatWords));
        CUDA_TRY(cudaMemsetAsync(ctx->d_tstats.p, 0, (size_t)n * 4 * kTaskStatWords, ctx->stream));
    }
    ctx->last_batch_n = n;
    ScanOut so;
    so.summary = ctx->d_summary.as<ScanSummary>();
    so.full_list = ctx->d_full_list.as<uint32_t>();
    so.ties = ctx->d_ties.as<TieEntry>();
    so.tie_cap = tie_cap;
    CUDA_TRY(cudaMemsetAsync(ctx->d_counters.as<unsigned long long>() + 7, 0, 8, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    CUDA_TRY(launch_scan(P, ctx->d_pb.as<uint16_t>(), n, n_plain, flags, ctx->d_records.as<dto_b200_record>(), so,
                         ctx->d_counters.as<unsigned long long>(),
                         ctx->opt_task_stats ? ctx->d_tstats.as<uint32_t>() : nullptr, grid, warps, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    count_launches(ctx, 1);
    ctx->stats.last_scan_launches += 1;
    ScanSummary *sum = ctx->h_summary.as<ScanSummary>();
    CUDA_TRY(cudaMemcpyAsync(sum, ctx->d_summary.p, sizeof(ScanSummary), cudaMemcpyDeviceToHost, ctx->stream));
    count_d2h(ctx, sizeof(ScanSummary));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.last_scan_kernel_ms += ms;
    if (sum->n_done != (uint32_t)n)
        return fail(DTO_B200_ERR_CUDA, "scan kernel finished %u of %d tasks", sum->n_done, n);
    const uint32_t n_full = sum->n_full;
    const uint32_t n_ties = std::min(sum->n_tie_entries, tie_cap);  // reservations past the cap were redirected to full_list
    ctx->stats.tasks_fast += (uint64_t)(n - (int)n_full);

    if (n_ties) {
        // entries of one task were reserved as one block, so they are contiguous
        CUDA_TRY(ctx->h_ties.ensure((size_t)n_ties * sizeof(TieEntry)));
        TieEntry *te = ctx->h_ties.as<TieEntry>();
        CUDA_TRY(cudaMemcpyAsync(te, ctx->d_ties.p, (size_t)n_ties * sizeof(TieEntry), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        count_d2h(ctx, (uint64_t)n_ties * sizeof(TieEntry));
        std::vector<uint32_t> idx;
        std::vector<dto_b200_record> rec;
        std::vector<HostCell> hc;
        for (uint32_t x = 0; x < n_ties;) {
            uint32_t y = x;
            hc.clear();
            if (te[x].task == 0xFFFFFFFFu) {  // void part of a reservation that straddled the end of the pool
                ++x;
                continue;
            }
            while (y < n_ties && te[y].task == te[x].task) {
                hc.push_back({te[y].ij, te[y].k});
                ++y;
            }
            if (te[x].task >= (uint32_t)n) return fail(DTO_B200_ERR_CUDA, "tie pool entry %u names task %u of %d", x, te[x].task, n);
            const uint32_t tf = (int)te[x].task < n_plain ? (flags & ~DTO_B200_FLAG_PERMUTED) : flags;
            idx.push_back(te[x].task);
            rec.push_back(resolve_on_host(ctx, hc.data(), hc.size(), tf | DTO_B200_FLAG_TIE_RESOLVED));
            x = y;
        }
        ctx->stats.tasks_tie_resolved += idx.size();
        const size_t m = idx.size();
        CUDA_TRY(ctx->h_patch.ensure(m * (sizeof(dto_b200_record) + 4)));
        CUDA_TRY(ctx->d_patch_idx.ensure(m * 4));
        CUDA_TRY(ctx->d_patch_rec.ensure(m * sizeof(dto_b200_record)));
        dto_b200_record *hp = ctx->h_patch.as<dto_b200_record>();
        uint32_t *hi = reinterpret_cast<uint32_t *>(hp + m);
        memcpy(hp, rec.data(), m * sizeof(dto_b200_record));
        memcpy(hi, idx.data(), m * 4);
        CUDA_TRY(cudaMemcpyAsync(ctx->d_patch_rec.p, hp, m * sizeof(dto_b200_record), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_patch_idx.p, hi, m * 4, cudaMemcpyHostToDevice, ctx->stream));
        count_h2d(ctx, m * (sizeof(dto_b200_record) + 4));
        CUDA_TRY(launch_patch_records(ctx->d_patch_idx.as<uint32_t>(), ctx->d_patch_rec.as<dto_b200_record>(), (int)m,
                                      ctx->d_records.as<dto_b200_record>(), ctx->stream));
        count_launches(ctx, 1);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the pinned staging buffers are reused by the next launch
    }
    if (n_full) {
        std::vector<uint32_t> full(n_full);
        CUDA_TRY(cudaMemcpy(full.data(), ctx->d_full_list.p, (size_t)n_full * 4, cudaMemcpyDeviceToHost));
        count_d2h(ctx, (uint64_t)n_full * 4);
        for (uint32_t t : full) {
            if (t >= (uint32_t)n) return fail(DTO_B200_ERR_CUDA, "dense-path list names task %u of %d", t, n);
            int rc = dense_task(ctx, ctx->d_pb.as<uint16_t>() + (size_t)t * P.pb_stride,
                                (int)t < n_plain ? (flags & ~DTO_B200_FLAG_PERMUTED) : flags,
                                ctx->d_records.as<dto_b200_record>() + t, nullptr);
            if (rc) return rc;
        }
    }
    return DTO_B200_OK;
}

int need_problem(dto_b200_ctx *ctx) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ctx->has_problem) return fail(DTO_B200_ERR_STATE, "no problem loaded: call dto_b200_set_problem first");
    return DTO_B200_OK;
}

int pull_counters(dto_b200_ctx *ctx) {
    unsigned long long c[8];
    CUDA_TRY(cudaMemcpy(c, ctx->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost));
    ctx->stats.candidates = c[0];
    ctx->stats.level2_cells = c[1];
    ctx->stats.refined_cells = c[2];
    return DTO_B200_OK;
}

}  // namespace

namespace dto {
void trim_batch_buffers(dto_b200_ctx *ctx, size_t keep_bytes) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_pb, &ctx->d_records, &ctx->d_full_list, &ctx->d_ties, &ctx->d_perm1, &ctx->d_perm2, &ctx->d_inv,
                      &ctx->d_words, &ctx->d_tstats, &ctx->d_minp, &ctx->d_slotmaps, &ctx->d_seeds};
    for (DevBuf *b : bufs)
        if (b->cap > keep_bytes) b->release();
}


// Batched list pairs (BASELINE config 4).  The problem loaded in ctx must have identical gene sets; the G pairs share its
// rank structure (hence its screen tables and pairing inputs) and differ in which gene sits where: pair g brings its own
// gene map slot_maps[g * n1 ..] (used by its unpermuted task only -- the null of identical gene sets does not depend on
// the map) and its own Philox seed.  One launch of each kernel covers the G unpermuted tasks (riding through the scan
// kernel as tasks 0 .. G-1) and the G x perms permuted ones (Philox ids 1 .. perms under seeds[g]).
// records_out is pair-major: pair g occupies [g * (perms + 1), (g + 1) * (perms + 1)), unpermuted record first.
int run_pair_group(dto_b200_ctx *ctx, const int32_t *slot_maps, const uint64_t *seeds, size_t G, size_t perms,
                   dto_b200_record *records_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (G == 0) return DTO_B200_OK;
    if (!slot_maps || !seeds || !records_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    const Problem &P = ctx->P;
    if (!(P.n_common == P.n1 && P.n_common == P.n2))
        return fail(DTO_B200_ERR_UNSUPPORTED, "pair groups need identical gene sets in both lists");
    const size_t n = G * (perms + 1);
    if (n > (size_t)1 << 24) return fail(DTO_B200_ERR_INVALID, "pair group too large (%zu tasks)", n);
    const int B1 = pick_bucket_bits(P.n1), B2 = pick_bucket_bits(P.n2);
    if (sigma_smem_bytes(P, B1, B2, false) > ctx->smem_optin)
        return fail(DTO_B200_ERR_UNSUPPORTED, "pairing kernel does not fit shared memory");
    const int sigma_grid_max = ctx->sm_count * 4;
    CUDA_TRY(ctx->d_words.ensure((size_t)sigma_grid_max * sigma_scratch_words(P) * 4));
    CUDA_TRY(ctx->d_pb.ensure(n * P.pb_stride * 2));
    CUDA_TRY(ctx->d_slotmaps.ensure(G * P.n1 * 4));
    CUDA_TRY(ctx->d_seeds.ensure(G * 8));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_slotmaps.p, slot_maps, G * P.n1 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_seeds.p, seeds, G * 8, cudaMemcpyHostToDevice, ctx->stream));
    count_h2d(ctx, G * ((uint64_t)P.n1 * 4 + 8));
    ctx->stats.last_scan_kernel_ms = 0;
    ctx->stats.last_sigma_kernel_ms = 0;
    ctx->stats.last_scan_launches = 0;
    CUDA_TRY(cudaEventRecord(ctx->ev[4], ctx->stream));
    CUDA_TRY(launch_compose(P, nullptr, nullptr, ctx->d_slotmaps.as<int32_t>(), (int)G, nullptr, ctx->d_err.as<int>(),
                            ctx->d_pb.as<uint16_t>(), ctx->stream));
    count_launches(ctx, 1);
    if (perms) {
        CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
        CUDA_TRY(launch_sigma_sort(P, 0, ctx->d_seeds.as<uint64_t>(), (uint32_t)perms, 1, (int)(G * perms),
                                   ctx->d_pb.as<uint16_t>() + G * P.pb_stride, nullptr, ctx->d_words.as<uint32_t>(),
                                   ctx->smem_optin, ctx->sm_count, ctx->opt_sigma_ctas, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
        count_launches(ctx, 1);
    }
    rc = run_tasks(ctx, (int)n, DTO_B200_FLAG_PERMUTED, (int)G);
    if (rc) return rc;
    if (perms) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        ctx->stats.last_sigma_kernel_ms += ms;
    }
    CUDA_TRY(ctx->h_records.ensure(n * sizeof(dto_b200_record)));
    dto_b200_record *stage = ctx->h_records.as<dto_b200_record>();
    CUDA_TRY(cudaMemcpyAsync(stage, ctx->d_records.p, n * sizeof(dto_b200_record), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev[5], ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    count_d2h(ctx, n * sizeof(dto_b200_record));
    float run_ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&run_ms, ctx->ev[4], ctx->ev[5]));
    ctx->stats.last_run_ms = run_ms;
    for (size_t g = 0; g < G; ++g) {
        dto_b200_record *dst = records_out + g * (perms + 1);
        dto_b200_record u = stage[g];
        if (!(u.flags & DTO_B200_FLAG_HOST_PVALUE)) {
            // the unpermuted p is an output field of the run (main.rs:165): evaluate it with the host libm, as the
            // reference does (the cell itself was settled exactly by the scan / tie resolution)
            u.pvalue = host_p(ctx, u.set1_len, u.set2_len, u.intersection_size);
            u.flags |= DTO_B200_FLAG_HOST_PVALUE;
        }
        dst[0] = u;
        memcpy(dst + 1, stage + G + g * perms, perms * sizeof(dto_b200_record));
    }
    return DTO_B200_OK;
}

// The pairing inputs of a problem: what makes two list pairs batchable into one group (together with identical gene sets)
bool same_rank_structure(const dto_b200_ctx *ctx, const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                         const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2, uint64_t population) {
    if (!ctx->has_problem) return false;
    const Problem &P = ctx->P;
    if (P.N != population || P.n1 != n1 || P.n2 != n2 || (size_t)P.T1 != T1 || (size_t)P.T2 != T2) return false;
    if (memcmp(ctx->h_thr1.data(), thr1, T1 * 4) || memcmp(ctx->h_thr2.data(), thr2, T2 * 4)) return false;
    return ctx->h_ranks1.size() == n1 && ctx->h_ranks2.size() == n2 && !memcmp(ctx->h_ranks1.data(), ranks1, n1 * 4) &&
           !memcmp(ctx->h_ranks2.data(), ranks2, n2 * 4);
}

bool identical_gene_sets(const dto_b200_ctx *ctx) {
    return ctx->has_problem && ctx->P.n_common == ctx->P.n1 && ctx->P.n_common == ctx->P.n2;
}

size_t group_task_capacity(const dto_b200_ctx *ctx) { return ctx->has_problem ? (size_t)auto_batch(ctx) : 0; }

}  // namespace dto

extern "C" {

const char *dto_b200_last_error(void) { return g_last_error.c_str(); }
const char *dto_b200_version(void) { return "dto-b200 0.2.0 (sm_100a)"; }

int dto_b200_device_count(int *count_out) {
    if (!count_out) return fail(DTO_B200_ERR_INVALID, "null count_out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count_out = 0;
        return fail(DTO_B200_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count_out = n;
    return DTO_B200_OK;
}

int dto_b200_create(dto_b200_ctx **ctx_out, int device) {
    if (!ctx_out) return fail(DTO_B200_ERR_INVALID, "null ctx_out");
    *ctx_out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(DTO_B200_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(DTO_B200_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(DTO_B200_ERR_CUDA, "device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major,
                    prop.minor);
    dto_b200_ctx *ctx = new dto_b200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(DTO_B200_ERR_CUDA, "cudaStreamCreate failed");
    }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    if (ctx->d_counters.ensure(64) != cudaSuccess || ctx->d_err.ensure(16) != cudaSuccess) {
        dto_b200_destroy(ctx);
        return fail(DTO_B200_ERR_CUDA, "cudaMalloc failed");
    }
    cudaMemset(ctx->d_counters.p, 0, 64);
    *ctx_out = ctx;
    return DTO_B200_OK;
}

void dto_b200_destroy(dto_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_c1, &ctx->d_c2, &ctx->d_thr1, &ctx->d_thr2, &ctx->d_lf, &ctx->d_rowA, &ctx->d_colB,
                      &ctx->d_kcrit, &ctx->d_meta, &ctx->d_lptab, &ctx->d_counts, &ctx->d_dslot2, &ctx->d_bin1, &ctx->d_bin2, &ctx->d_slot2, &ctx->d_pb,
                      &ctx->d_records, &ctx->d_counters, &ctx->d_H,
                      &ctx->d_pv, &ctx->d_logp, &ctx->d_perm1, &ctx->d_perm2, &ctx->d_inv, &ctx->d_err, &ctx->d_pair,
                      &ctx->d_minp, &ctx->d_tstats, &ctx->d_words, &ctx->d_summary, &ctx->d_full_list, &ctx->d_ties,
                      &ctx->d_patch_idx, &ctx->d_patch_rec, &ctx->d_best_cell, &ctx->d_collect, &ctx->d_slotmaps, &ctx->d_seeds};
    for (DevBuf *b : bufs) b->release();
    ctx->h_records.release();
    ctx->h_stage.release();
    ctx->h_summary.release();
    ctx->h_ties.release();
    ctx->h_patch.release();
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int dto_b200_set_option(dto_b200_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return fail(DTO_B200_ERR_INVALID, "null argument");
    const std::string s(name);
    if (s == "batch") {
        if (value < 0 || value > (1 << 22)) return fail(DTO_B200_ERR_INVALID, "batch out of range");
        ctx->opt_batch = (int)value;
    } else if (s == "warps_per_cta") {
        if (value < 1 || value > kScanThreads / 32) return fail(DTO_B200_ERR_INVALID, "warps_per_cta out of range");
        ctx->opt_warps = (int)value;
    } else if (s == "packed_screen") {
        if (ctx->has_problem) return fail(DTO_B200_ERR_STATE, "set 'packed_screen' before dto_b200_set_problem");
        ctx->opt_swar = value != 0;
    } else if (s == "table_cache") {
        ctx->opt_table_cache = value != 0;
    } else if (s == "sigma_ctas") {
        if (value < 0 || value > 2) return fail(DTO_B200_ERR_INVALID, "sigma_ctas must be 0, 1 or 2");
        ctx->opt_sigma_ctas = (int)value;
    } else if (s == "task_stats") {
        ctx->opt_task_stats = value != 0;
    } else if (s == "levels") {
        // fewer than two levels leave no tabulated range (tau_1 .. tau_levels), and level 0 alone certifies nothing
        if (value < 2 || value > kMaxLevels) return fail(DTO_B200_ERR_INVALID, "levels must be 2..%d", kMaxLevels);
        if (ctx->has_problem) return fail(DTO_B200_ERR_STATE, "set 'levels' before dto_b200_set_problem");
        ctx->opt_levels = (int)value;
    } else {
        return fail(DTO_B200_ERR_INVALID, "unknown option '%s'", name);
    }
    return DTO_B200_OK;
}

int dto_b200_get_limits(dto_b200_limits *out) {
    if (!out) return fail(DTO_B200_ERR_INVALID, "null output");
    out->max_features_per_list = 65534;
    out->max_thresholds_per_list = 2048;
    out->max_population = (uint64_t)1 << 27;
    out->max_tasks_per_call = (uint64_t)1 << 40;
    return DTO_B200_OK;
}

int dto_b200_process_totals(uint64_t *out3) {
    if (!out3) return fail(DTO_B200_ERR_INVALID, "null output");
    out3[0] = g_total_launches.load();
    out3[1] = g_total_h2d.load();
    out3[2] = g_total_d2h.load();
    return DTO_B200_OK;
}

int dto_b200_get_stats(dto_b200_ctx *ctx, dto_b200_stats *out) {
    if (!ctx || !out) return fail(DTO_B200_ERR_INVALID, "null argument");
    int rc = bind(ctx);
    if (rc) return rc;
    rc = pull_counters(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return DTO_B200_OK;
}

int dto_b200_last_batch_task_stats(dto_b200_ctx *ctx, uint32_t *out, size_t max_tasks, size_t *n_out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!out || !n_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    if (!ctx->opt_task_stats) return fail(DTO_B200_ERR_STATE, "set option 'task_stats' to 1 first");
    const size_t n = std::min(max_tasks, (size_t)ctx->last_batch_n);
    CUDA_TRY(cudaMemcpy(out, ctx->d_tstats.p, n * 4 * kTaskStatWords, cudaMemcpyDeviceToHost));
    *n_out = n;
    return DTO_B200_OK;
}

int dto_b200_table_logp(dto_b200_ctx *ctx, const uint32_t *row, const uint32_t *col, const uint32_t *k, size_t count,
                        double *logp_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (count == 0) return DTO_B200_OK;
    if (!row || !col || !k || !logp_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    const Problem &P = ctx->P;
    for (size_t x = 0; x < count; ++x) {
        logp_out[x] = NAN;
        if (row[x] >= (uint32_t)P.T1 || col[x] >= (uint32_t)P.T2) return fail(DTO_B200_ERR_INVALID, "cell %zu out of range", x);
        uint2 meta;
        CUDA_TRY(cudaMemcpy(&meta, P.cellmeta + (size_t)row[x] * P.T2 + col[x], sizeof(meta), cudaMemcpyDeviceToHost));
        const uint32_t kbase = meta.y & 0xFFFFu, cnt = meta.y >> 16;
        if (k[x] >= kbase && k[x] - kbase < cnt)
            CUDA_TRY(cudaMemcpy(&logp_out[x], P.lptab + meta.x + (k[x] - kbase), sizeof(double), cudaMemcpyDeviceToHost));
    }
    return DTO_B200_OK;
}

int dto_b200_reset_stats(dto_b200_ctx *ctx) {
    if (!ctx) return fail(DTO_B200_ERR_INVALID, "null context");
    int rc = bind(ctx);
    if (rc) return rc;
    ctx->stats = dto_b200_stats{};
    CUDA_TRY(cudaMemset(ctx->d_counters.p, 0, 64));
    return DTO_B200_OK;
}

int dto_b200_set_problem(dto_b200_ctx *ctx, const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                         const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                         const int32_t *slot2_of_1, uint64_t population) {
    int rc = bind(ctx);
    if (rc) return rc;
    ctx->has_problem = false;
    if ((n1 && !ranks1) || (n2 && !ranks2) || (T1 && !thr1) || (T2 && !thr2) || (n1 && !slot2_of_1))
        return fail(DTO_B200_ERR_INVALID, "null input array");
    if (T1 == 0 || T2 == 0)
        return fail(DTO_B200_ERR_PANIC,
                    "called `Option::unwrap()` on a `None` value: a ranked list has no thresholds (empty list or "
                    "max rank 0; optimize_main.rs:116)");
    if (n1 > 65534 || n2 > 65534)
        return fail(DTO_B200_ERR_UNSUPPORTED, "lists longer than 65534 features are not supported (n1=%zu n2=%zu)", n1, n2);
    if (T1 > 2048 || T2 > 2048) return fail(DTO_B200_ERR_UNSUPPORTED, "more than 2048 thresholds per list");
    if (population > ((uint64_t)1 << 27))
        return fail(DTO_B200_ERR_UNSUPPORTED, "population %llu exceeds the ln-factorial table limit (2^27)",
                    (unsigned long long)population);
    for (size_t j = 1; j < n1; ++j)
        if (ranks1[j] < ranks1[j - 1]) return fail(DTO_B200_ERR_INVALID, "ranks1 must be sorted ascending");
    for (size_t j = 1; j < n2; ++j)
        if (ranks2[j] < ranks2[j - 1]) return fail(DTO_B200_ERR_INVALID, "ranks2 must be sorted ascending");
    for (size_t i = 1; i < T1; ++i)
        if (thr1[i] <= thr1[i - 1]) return fail(DTO_B200_ERR_INVALID, "thresholds1 must be strictly increasing");
    for (size_t i = 1; i < T2; ++i)
        if (thr2[i] <= thr2[i - 1]) return fail(DTO_B200_ERR_INVALID, "thresholds2 must be strictly increasing");
    std::vector<uint8_t> seen(n2, 0);
    uint32_t n_common = 0;
    for (size_t a = 0; a < n1; ++a) {
        const int32_t s = slot2_of_1[a];
        if (s < 0) continue;
        if ((size_t)s >= n2) return fail(DTO_B200_ERR_INVALID, "slot2_of_1[%zu]=%d out of range", a, s);
        if (seen[s]) return fail(DTO_B200_ERR_INVALID, "slot2_of_1 is not injective (list-2 slot %d used twice)", s);
        seen[s] = 1;
        ++n_common;
    }

    Problem P{};
    P.T1 = (int)T1;
    P.T2 = (int)T2;
    P.CH = pick_ch((int)T2);
    P.CHP = hist_words(P.CH);
    P.T2pad = kcrit_row_u16(P.CH);
    P.levels = ctx->opt_levels;
    P.n1 = (uint32_t)n1;
    P.n2 = (uint32_t)n2;
    P.N = population;
    P.n_common = n_common;
    std::vector<uint32_t> c1(T1), c2(T2);
    for (size_t i = 0; i < T1; ++i) c1[i] = (uint32_t)(std::upper_bound(ranks1, ranks1 + n1, thr1[i]) - ranks1);
    for (size_t i = 0; i < T2; ++i) c2[i] = (uint32_t)(std::upper_bound(ranks2, ranks2 + n2, thr2[i]) - ranks2);
    if (c1[T1 - 1] > population || c2[T2 - 1] > population)
        return fail(DTO_B200_ERR_PANIC,
                    "Failed to create hypergeometric distribution: a feature set (%u / %u) is larger than the population "
                    "(%llu)",
                    c1[T1 - 1], c2[T2 - 1], (unsigned long long)population);
    P.n1_eff = c1[T1 - 1];
    P.never = (ctx->opt_swar && c1[T1 - 1] <= 32766u && c2[T2 - 1] <= 32766u) ? 0x7FFFu : 0xFFFFu;
    P.pb_stride = ((P.n1_eff + 1 + 255) / 256) * 256;  // whole staging chunks (cp.async in the scan), 512 B aligned rows
    std::vector<uint16_t> bin1(n1 ? n1 : 1), bin2(n2 ? n2 : 1), dslot2(((n2 ? n2 : 1) + 7) & ~(size_t)7, kNoSlot);
    for (size_t j = 0; j < n1; ++j) {
        const size_t b = std::lower_bound(thr1, thr1 + T1, ranks1[j]) - thr1;
        bin1[j] = b < T1 ? (uint16_t)b : kNoSlot;
    }
    for (size_t j = 0; j < n2; ++j) {
        const size_t b = std::lower_bound(thr2, thr2 + T2, ranks2[j]) - thr2;
        bin2[j] = b < T2 ? (uint16_t)b : kNoSlot;
        dslot2[j] = b < T2 ? (uint16_t)hist_slot((int)b, P.CH) : kNoSlot;
    }
    const bool lf_cached = (ctx->lf_N == population && ctx->lf_host.size() == population + 1 && ctx->d_lf.p != nullptr);
    if (!lf_cached) {
        host_fill_ln_factorial(ctx->lf_host, population);
        ctx->lf_N = ~0ull;  // set after the upload below succeeds
    }
    const std::vector<double> &lf = ctx->lf_host;
    std::vector<double> rowA(T1), colB(T2);
    for (size_t i = 0; i < T1; ++i) rowA[i] = lf[c1[i]] + lf[population - c1[i]];
    for (size_t j = 0; j < T2; ++j) colB[j] = lf[c2[j]] + lf[population - c2[j]] - lf[population];
    {
        // error budget of log pmf (a 7-term sum of ln-factorials of magnitude <= lf[N]): see Problem::refine_eps
        const double top = lf[population] > 1.0 ? lf[population] : 1.0;
        const double ulp = std::nextafter(top, INFINITY) - top;
        P.refine_eps = std::max(kRefineEps, 40.0 * ulp);
        P.tab_slack = std::max(1e-6, 4.0 * P.refine_eps);
    }
    P.level_log[0] = INFINITY;
    // tau_1 = 0.95 ends the unscreened phase at once; then half-octave steps 0.5, 0.354, 0.25, ... so that a permutation
    // whose running minimum sits just above a level still screens out everything more than ~1.4x above it
    P.level_log[1] = std::log(0.95);
    for (int l = 2; l <= kMaxLevels; ++l) P.level_log[l] = std::log(0.5) - 0.5 * (double)(l - 2) * std::log(2.0);

    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.ensure(bytes);
        if (e != cudaSuccess) return e;
        count_h2d(ctx, bytes);
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    ctx->h_c1 = c1;
    ctx->h_c2 = c2;
    ctx->h_thr1.assign(thr1, thr1 + T1);
    ctx->h_thr2.assign(thr2, thr2 + T2);
    ctx->h_ranks1.assign(ranks1, ranks1 + n1);
    ctx->h_ranks2.assign(ranks2, ranks2 + n2);
    const bool tables_cached = ctx->opt_table_cache && ctx->tab_valid && ctx->tab_N == population &&
                               ctx->tab_levels == P.levels && ctx->tab_never == P.never && ctx->tab_c1 == c1 &&
                               ctx->tab_c2 == c2;
    if (!tables_cached) {
        ctx->tab_valid = false;
        CUDA_TRY(up(ctx->d_c1, c1.data(), T1 * 4));
        CUDA_TRY(up(ctx->d_c2, c2.data(), T2 * 4));
        CUDA_TRY(up(ctx->d_rowA, rowA.data(), T1 * 8));
        CUDA_TRY(up(ctx->d_colB, colB.data(), T2 * 8));
    }
    CUDA_TRY(up(ctx->d_thr1, thr1, T1 * 4));
    CUDA_TRY(up(ctx->d_thr2, thr2, T2 * 4));
    if (!lf_cached) {
        CUDA_TRY(up(ctx->d_lf, lf.data(), lf.size() * 8));
        ctx->lf_N = population;
    }
    CUDA_TRY(up(ctx->d_dslot2, dslot2.data(), dslot2.size() * 2));
    CUDA_TRY(up(ctx->d_bin1, bin1.data(), bin1.size() * 2));
    CUDA_TRY(up(ctx->d_bin2, bin2.data(), bin2.size() * 2));
    CUDA_TRY(up(ctx->d_slot2, slot2_of_1, n1 * 4));  // n1 >= 1 here (T1 >= 1)
    CUDA_TRY(ctx->d_kcrit.ensure((size_t)(P.levels + 1) * T1 * P.T2pad * 2));
    P.c1 = ctx->d_c1.as<uint32_t>();
    P.c2 = ctx->d_c2.as<uint32_t>();
    P.thr1 = ctx->d_thr1.as<uint32_t>();
    P.thr2 = ctx->d_thr2.as<uint32_t>();
    P.lf = ctx->d_lf.as<double>();
    P.rowA = ctx->d_rowA.as<double>();
    P.colB = ctx->d_colB.as<double>();
    P.kcrit = ctx->d_kcrit.as<uint16_t>();
    P.dslot2 = ctx->d_dslot2.as<uint16_t>();
    P.bin1 = ctx->d_bin1.as<uint16_t>();
    P.bin2 = ctx->d_bin2.as<uint16_t>();
    P.slot2_of_1 = ctx->d_slot2.as<int32_t>();
    P.cellmeta = ctx->d_meta.as<uint2>();
    P.lptab = ctx->d_lptab.as<double>();
    if (tables_cached) {
        ++ctx->stats.table_cache_hits;
        ctx->stats.lptab_entries = ctx->tab_entries;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
        ctx->P = P;
        ctx->has_problem = true;
        return DTO_B200_OK;
    }
    // screen tables, then the log-p table: counts -> exclusive scan (host; a few hundred thousand words) -> fill
    const size_t cells = T1 * T2;
    CUDA_TRY(ctx->d_counts.ensure(cells * 4));
    CUDA_TRY(ctx->d_meta.ensure(cells * 8));
    CUDA_TRY(launch_build_kcrit(P, ctx->d_kcrit.as<uint16_t>(), ctx->d_counts.as<uint32_t>(), ctx->d_meta.as<uint2>(), ctx->stream));
    // offsets of the per-cell table segments = exclusive scan of the counts, on the device; only the total comes back
    // (it sizes the table)
    unsigned long long *d_total = ctx->d_counters.as<unsigned long long>() + 6;
    CUDA_TRY(launch_scan_counts(ctx->d_counts.as<uint32_t>(), (int)cells, d_total, ctx->stream));
    unsigned long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (total > 0xFFFFFFFFull) return fail(DTO_B200_ERR_UNSUPPORTED, "log-p table too large (%llu entries)", total);
    CUDA_TRY(ctx->d_lptab.ensure((size_t)(total ? total : 1) * 8));
    P.cellmeta = ctx->d_meta.as<uint2>();
    P.lptab = ctx->d_lptab.as<double>();
    CUDA_TRY(launch_fill_lptab(P, ctx->d_counts.as<uint32_t>(), ctx->d_meta.as<uint2>(), ctx->d_lptab.as<double>(), ctx->stream));
    count_launches(ctx, 4);
    count_d2h(ctx, 8);
    ctx->stats.lptab_entries = total;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    ctx->tab_N = population;
    ctx->tab_levels = P.levels;
    ctx->tab_never = P.never;
    ctx->tab_entries = total;
    ctx->tab_c1.swap(c1);
    ctx->tab_c2.swap(c2);
    ctx->tab_valid = true;
    ctx->P = P;
    ctx->has_problem = true;
    return DTO_B200_OK;
}

int dto_b200_run_unpermuted(dto_b200_ctx *ctx, dto_b200_record *record_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (!record_out) return fail(DTO_B200_ERR_INVALID, "null record_out");
    const Problem &P = ctx->P;
    CUDA_TRY(ctx->d_pb.ensure((size_t)P.pb_stride * 2));
    CUDA_TRY(launch_compose(P, nullptr, nullptr, nullptr, 1, nullptr, ctx->d_err.as<int>(), ctx->d_pb.as<uint16_t>(), ctx->stream));
    count_launches(ctx, 1);
    ctx->stats.last_scan_kernel_ms = 0;
    ctx->stats.last_sigma_kernel_ms = 0;
    ctx->stats.last_scan_launches = 0;
    // One task cannot fill the GPU through the warp-per-permutation scan, and for strongly concordant lists nearly
    // every cell is deep in the tail; the dense grid (every cell evaluated in statrs order on all SMs) is both exact
    // and fast here: optimize(l1, l2, permute = false, ..) == argmin of the debug grid (optimize_main.rs:73-116).
    // The optimum itself is settled on the host (host libm), so the record's p is bit-identical to the reference's.
    return dense_task(ctx, ctx->d_pb.as<uint16_t>(), 0u, nullptr, record_out);
}

int dto_b200_run_permuted_indices(dto_b200_ctx *ctx, const uint32_t *perm1, const uint32_t *perm2, size_t Pn,
                                  dto_b200_record *records_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (Pn == 0) return DTO_B200_OK;
    if (!perm1 || !perm2 || !records_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    const Problem &P = ctx->P;
    // bound the staging footprint to ~512 MiB of indices per batch
    const size_t per_task = ((size_t)P.n1 + P.n2 + std::max(P.n1, P.n2)) * 4 + (size_t)P.pb_stride * 2;
    size_t batch = std::max<size_t>(1, ((size_t)512 << 20) / per_task);
    batch = std::min(batch, (size_t)auto_batch(ctx));
    ctx->stats.last_scan_kernel_ms = 0;
    ctx->stats.last_sigma_kernel_ms = 0;
    ctx->stats.last_scan_launches = 0;
    for (size_t done = 0; done < Pn; done += batch) {
        const int n = (int)std::min(batch, Pn - done);
        CUDA_TRY(ctx->d_perm1.ensure((size_t)n * P.n1 * 4));
        CUDA_TRY(ctx->d_perm2.ensure((size_t)n * P.n2 * 4));
        CUDA_TRY(ctx->d_inv.ensure((size_t)n * std::max(P.n1, P.n2) * 4));
        CUDA_TRY(ctx->d_pb.ensure((size_t)n * P.pb_stride * 2));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_perm1.p, perm1 + done * P.n1, (size_t)n * P.n1 * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_perm2.p, perm2 + done * P.n2, (size_t)n * P.n2 * 4, cudaMemcpyHostToDevice, ctx->stream));
        count_h2d(ctx, (uint64_t)n * ((uint64_t)P.n1 + P.n2) * 4);
        CUDA_TRY(cudaMemsetAsync(ctx->d_err.p, 0, 4, ctx->stream));
        CUDA_TRY(launch_compose(P, ctx->d_perm1.as<uint32_t>(), ctx->d_perm2.as<uint32_t>(), nullptr, n, ctx->d_inv.as<uint32_t>(),
                                ctx->d_err.as<int>(), ctx->d_pb.as<uint16_t>(), ctx->stream));
        count_launches(ctx, 5);
        int err = 0;
        CUDA_TRY(cudaMemcpyAsync(&err, ctx->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (err) return fail(DTO_B200_ERR_INVALID, "perm1/perm2 rows must each be a permutation of 0..n-1");
        rc = run_tasks(ctx, n, DTO_B200_FLAG_PERMUTED);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpy(records_out + done, ctx->d_records.p, (size_t)n * sizeof(dto_b200_record), cudaMemcpyDeviceToHost));
        count_d2h(ctx, (uint64_t)n * sizeof(dto_b200_record) + 4);
    }
    return DTO_B200_OK;
}

static int run_philox_common(dto_b200_ctx *ctx, uint64_t seed, uint64_t first, size_t Pn, dto_b200_record *h_records,
                             double *h_minp, dto_b200_record *d_records_out, double *d_minp_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (Pn == 0) return DTO_B200_OK;
    const Problem &P = ctx->P;
    const int B1 = pick_bucket_bits(P.n1), B2 = pick_bucket_bits(P.n2);
    // The sort buffer, secondary keys and boundary list of the pairing kernel live in a per-CTA global scratch (L2-
    // resident): the row-wise fast path touches it for ~5 % of the elements only, and long lists would not fit it in
    // shared memory anyway.
    if (sigma_smem_bytes(P, B1, B2, false) > ctx->smem_optin)
        return fail(DTO_B200_ERR_UNSUPPORTED, "pairing kernel needs %zu B of shared memory (> %zu)",
                    sigma_smem_bytes(P, B1, B2, false), ctx->smem_optin);
    const int sigma_grid_max = ctx->sm_count * 4;
    CUDA_TRY(ctx->d_words.ensure((size_t)sigma_grid_max * sigma_scratch_words(P) * 4));
    uint32_t *words_scratch = ctx->d_words.as<uint32_t>();
    const size_t batch = (size_t)auto_batch(ctx);
    ctx->stats.last_scan_kernel_ms = 0;
    ctx->stats.last_sigma_kernel_ms = 0;
    ctx->stats.last_scan_launches = 0;
    if (h_records || h_minp) CUDA_TRY(ctx->h_records.ensure(std::min(batch, Pn) * sizeof(dto_b200_record)));
    cudaEvent_t run_begin = ctx->ev[4], run_end = ctx->ev[5];
    CUDA_TRY(cudaEventRecord(run_begin, ctx->stream));
    for (size_t done = 0; done < Pn; done += batch) {
        const int n = (int)std::min(batch, Pn - done);
        CUDA_TRY(ctx->d_pb.ensure((size_t)n * P.pb_stride * 2));
        CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
        CUDA_TRY(launch_sigma_sort(P, seed, nullptr, 0, first + done, n, ctx->d_pb.as<uint16_t>(), nullptr, words_scratch,
                                   ctx->smem_optin, ctx->sm_count, ctx->opt_sigma_ctas, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
        count_launches(ctx, 1);
        rc = run_tasks(ctx, n, DTO_B200_FLAG_PERMUTED);
        if (rc) return rc;
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        ctx->stats.last_sigma_kernel_ms += ms;
        if (h_records || h_minp) {
            dto_b200_record *stage = ctx->h_records.as<dto_b200_record>();
            CUDA_TRY(cudaMemcpyAsync(stage, ctx->d_records.p, (size_t)n * sizeof(dto_b200_record), cudaMemcpyDeviceToHost, ctx->stream));
            count_d2h(ctx, (uint64_t)n * sizeof(dto_b200_record));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (h_records) memcpy(h_records + done, stage, (size_t)n * sizeof(dto_b200_record));
            if (h_minp)
                for (int t = 0; t < n; ++t) h_minp[done + t] = stage[t].pvalue;
        }
        if (d_records_out)
            CUDA_TRY(cudaMemcpyAsync(d_records_out + done, ctx->d_records.p, (size_t)n * sizeof(dto_b200_record), cudaMemcpyDeviceToDevice, ctx->stream));
        if (d_minp_out)
            CUDA_TRY(cudaMemcpy2DAsync(d_minp_out + done, sizeof(double),
                                       reinterpret_cast<const char *>(ctx->d_records.p) + offsetof(dto_b200_record, pvalue),
                                       sizeof(dto_b200_record), sizeof(double), (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        // no host synchronisation here: the copies are ordered before the next batch's kernels on the same stream, and the
        // event synchronisation below covers the last batch
    }
    CUDA_TRY(cudaEventRecord(run_end, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(run_end));
    float run_ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&run_ms, run_begin, run_end));
    ctx->stats.last_run_ms = run_ms;
    return DTO_B200_OK;
}

int dto_b200_run_permuted_philox(dto_b200_ctx *ctx, uint64_t seed, uint64_t first_perm_id, size_t Pn,
                                 dto_b200_record *records_out, double *minp_out) {
    return run_philox_common(ctx, seed, first_perm_id, Pn, records_out, minp_out, nullptr, nullptr);
}

int dto_b200_run_permuted_philox_device(dto_b200_ctx *ctx, uint64_t seed, uint64_t first_perm_id, size_t Pn,
                                        double *d_minp_out, dto_b200_record *d_records_out) {
    return run_philox_common(ctx, seed, first_perm_id, Pn, nullptr, nullptr, d_records_out, d_minp_out);
}

int dto_b200_philox_pairing(dto_b200_ctx *ctx, uint64_t seed, uint64_t perm_id, uint32_t *pos2_of_pos1_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if (!pos2_of_pos1_out) return fail(DTO_B200_ERR_INVALID, "null output");
    const Problem &P = ctx->P;
    const int B1 = pick_bucket_bits(P.n1), B2 = pick_bucket_bits(P.n2);
    if (sigma_smem_bytes(P, B1, B2, false) > ctx->smem_optin)
        return fail(DTO_B200_ERR_UNSUPPORTED, "pairing kernel does not fit shared memory");
    CUDA_TRY(ctx->d_words.ensure(sigma_scratch_words(P) * 4));
    uint32_t *words_scratch = ctx->d_words.as<uint32_t>();
    CUDA_TRY(ctx->d_pb.ensure((size_t)P.pb_stride * 2));
    CUDA_TRY(ctx->d_pair.ensure((size_t)(P.n1 ? P.n1 : 1) * 4));
    CUDA_TRY(cudaMemsetAsync(ctx->d_pair.p, 0xFF, (size_t)(P.n1 ? P.n1 : 1) * 4, ctx->stream));
    CUDA_TRY(launch_sigma_sort(P, seed, nullptr, 0, perm_id, 1, ctx->d_pb.as<uint16_t>(), ctx->d_pair.as<uint32_t>(), words_scratch,
                               ctx->smem_optin, ctx->sm_count, 1, ctx->stream));
    count_launches(ctx, 1);
    CUDA_TRY(cudaMemcpyAsync(pos2_of_pos1_out, ctx->d_pair.p, (size_t)P.n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DTO_B200_OK;
}

int dto_b200_grid_debug(dto_b200_ctx *ctx, const uint32_t *perm1, const uint32_t *perm2, uint32_t *overlap_out,
                        double *pvalue_out, double *logp_out) {
    int rc = need_problem(ctx);
    if (rc) return rc;
    if ((perm1 == nullptr) != (perm2 == nullptr))
        return fail(DTO_B200_ERR_INVALID, "perm1 and perm2 must both be given or both be NULL");
    const Problem &P = ctx->P;
    const size_t cells = (size_t)P.T1 * P.T2;
    CUDA_TRY(ctx->d_pb.ensure((size_t)P.pb_stride * 2));
    CUDA_TRY(ctx->d_H.ensure(cells * 4));
    CUDA_TRY(ctx->d_pv.ensure(cells * 8));
    CUDA_TRY(ctx->d_logp.ensure(cells * 8));
    CUDA_TRY(cudaMemsetAsync(ctx->d_err.p, 0, 4, ctx->stream));
    if (perm1) {
        CUDA_TRY(ctx->d_perm1.ensure((size_t)P.n1 * 4));
        CUDA_TRY(ctx->d_perm2.ensure((size_t)P.n2 * 4));
        CUDA_TRY(ctx->d_inv.ensure((size_t)std::max(P.n1, P.n2) * 4));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_perm1.p, perm1, (size_t)P.n1 * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_perm2.p, perm2, (size_t)P.n2 * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(launch_compose(P, perm1 ? ctx->d_perm1.as<uint32_t>() : nullptr, perm2 ? ctx->d_perm2.as<uint32_t>() : nullptr,
                            nullptr, 1, ctx->d_inv.as<uint32_t>(), ctx->d_err.as<int>(), ctx->d_pb.as<uint16_t>(), ctx->stream));
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, ctx->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (err) return fail(DTO_B200_ERR_INVALID, "perm1/perm2 must each be a permutation of 0..n-1");
    CUDA_TRY(launch_full_grid(P, ctx->d_pb.as<uint16_t>(), ctx->d_H.as<uint32_t>(), pvalue_out ? ctx->d_pv.as<double>() : nullptr,
                              logp_out ? ctx->d_logp.as<double>() : nullptr, ctx->stream));
    count_launches(ctx, 5);
    if (overlap_out) CUDA_TRY(cudaMemcpyAsync(overlap_out, ctx->d_H.p, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (pvalue_out) CUDA_TRY(cudaMemcpyAsync(pvalue_out, ctx->d_pv.p, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (logp_out) CUDA_TRY(cudaMemcpyAsync(logp_out, ctx->d_logp.p, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DTO_B200_OK;
}

int dto_b200_hypergeometric_pvalues(dto_b200_ctx *ctx, const uint64_t *N, const uint64_t *K, const uint64_t *n,
                                    const uint64_t *k, size_t count, double *pvalue_out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (count == 0) return DTO_B200_OK;
    if (!N || !K || !n || !k || !pvalue_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    uint64_t maxN = 0;
    for (size_t x = 0; x < count; ++x) {
        maxN = std::max(maxN, N[x]);
        // Hypergeometric::new(N, K, n).expect(..) (hypergeometric_pvalue.rs:40-41) panics unless K <= N and n <= N
        if (K[x] > N[x] || n[x] > N[x])
            return fail(DTO_B200_ERR_PANIC,
                        "Failed to create hypergeometric distribution: quadruple %zu has successes %llu / draws %llu > population %llu",
                        x, (unsigned long long)K[x], (unsigned long long)n[x], (unsigned long long)N[x]);
    }
    if (maxN > ((uint64_t)1 << 27)) return fail(DTO_B200_ERR_UNSUPPORTED, "population exceeds 2^27");
    std::vector<double> lf;
    host_fill_ln_factorial(lf, maxN);
    DevBuf d_lf, d_in, d_out;
    CUDA_TRY(d_lf.ensure(lf.size() * 8));
    CUDA_TRY(d_in.ensure(count * 8 * 4));
    CUDA_TRY(d_out.ensure(count * 8));
    uint64_t *din = d_in.as<uint64_t>();
    CUDA_TRY(cudaMemcpy(d_lf.p, lf.data(), lf.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(din, N, count * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(din + count, K, count * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(din + 2 * count, n, count * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(din + 3 * count, k, count * 8, cudaMemcpyHostToDevice));
    cudaError_t e = launch_pvalues(d_lf.as<double>(), din, din + count, din + 2 * count, din + 3 * count, count,
                                   d_out.as<double>(), ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(pvalue_out, d_out.p, count * 8, cudaMemcpyDeviceToHost);
    d_lf.release();
    d_in.release();
    d_out.release();
    count_launches(ctx, 1);
    CUDA_TRY(e);
    return DTO_B200_OK;
}

int dto_b200_probe_fp64_tflops(dto_b200_ctx *ctx, double *tflops_out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!tflops_out) return fail(DTO_B200_ERR_INVALID, "null output");
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    DevBuf out;
    CUDA_TRY(out.ensure((size_t)blocks * threads * 8));
    CUDA_TRY(launch_fp64_probe(out.as<double>(), blocks, threads, 64, ctx->stream));  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
        CUDA_TRY(launch_fp64_probe(out.as<double>(), blocks, threads, iters, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        const double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    out.release();
    *tflops_out = best;
    return DTO_B200_OK;
}

int dto_b200_probe_hbm_gbs(dto_b200_ctx *ctx, double *gbs_out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!gbs_out) return fail(DTO_B200_ERR_INVALID, "null output");
    const size_t bytes = (size_t)1 << 30;
    DevBuf a, b;
    CUDA_TRY(a.ensure(bytes));
    CUDA_TRY(b.ensure(bytes));
    CUDA_TRY(cudaMemsetAsync(a.p, 1, bytes, ctx->stream));
    CUDA_TRY(launch_hbm_probe(a.p, b.p, bytes, ctx->sm_count * 16, ctx->stream));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
        CUDA_TRY(launch_hbm_probe(a.p, b.p, bytes, ctx->sm_count * 16, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        best = std::max(best, 2.0 * (double)bytes / (ms * 1e-3) / 1e9);
    }
    a.release();
    b.release();
    *gbs_out = best;
    return DTO_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Product-side collective: the all-gather of per-permutation minima that replaces the MPI gather of
// src/run/multi_node.rs:148-160 when one process drives each GPU.  NCCL is bound at run time (dlopen of libnccl.so.2:
// the copy already mapped by the host application if there is one, else the system's), so the library neither links
// NCCL nor needs it when a single process drives all GPUs.
// ---------------------------------------------------------------------------------------------------
namespace {

struct NcclUid {
    char internal[128];
};

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*CommCount)(void *, int *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};

NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) {
            api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(api.handle, name);
            if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

int nccl_ready(NcclApi *&out) {
    NcclApi &a = nccl_api();
    if (!a.why.empty() || !a.handle) return fail(DTO_B200_ERR_UNSUPPORTED, "NCCL unavailable: %s", a.why.c_str());
    out = &a;
    return DTO_B200_OK;
}

constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64 (nccl.h)

}  // namespace

extern "C" {

int dto_b200_nccl_unique_id(void *id_out) {
    if (!id_out) return fail(DTO_B200_ERR_INVALID, "null id_out");
    NcclApi *a = nullptr;
    int rc = nccl_ready(a);
    if (rc) return rc;
    const int e = a->GetUniqueId(reinterpret_cast<NcclUid *>(id_out));
    if (e) return fail(DTO_B200_ERR_CUDA, "ncclGetUniqueId: %s", a->GetErrorString(e));
    return DTO_B200_OK;
}

int dto_b200_nccl_comm_create(void **comm_out, int n_ranks, const void *unique_id, int rank, int device) {
    if (!comm_out || !unique_id) return fail(DTO_B200_ERR_INVALID, "null argument");
    *comm_out = nullptr;
    NcclApi *a = nullptr;
    int rc = nccl_ready(a);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    NcclUid uid;
    memcpy(&uid, unique_id, sizeof(uid));
    const int e = a->CommInitRank(comm_out, n_ranks, uid, rank);
    if (e) return fail(DTO_B200_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", rank, n_ranks, a->GetErrorString(e));
    return DTO_B200_OK;
}

int dto_b200_nccl_comm_destroy(void *comm) {
    if (!comm) return DTO_B200_OK;
    NcclApi *a = nullptr;
    int rc = nccl_ready(a);
    if (rc) return rc;
    const int e = a->CommDestroy(comm);
    if (e) return fail(DTO_B200_ERR_CUDA, "ncclCommDestroy: %s", a->GetErrorString(e));
    return DTO_B200_OK;
}

int dto_b200_allgather_minima(dto_b200_ctx *ctx, void *nccl_comm, const double *d_send, double *d_recv, size_t count) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!nccl_comm || (count && (!d_send || !d_recv))) return fail(DTO_B200_ERR_INVALID, "null argument");
    NcclApi *a = nullptr;
    rc = nccl_ready(a);
    if (rc) return rc;
    if (count == 0) return DTO_B200_OK;
    const int e = a->AllGather(d_send, d_recv, count, kNcclFloat64, nccl_comm, ctx->stream);
    if (e) return fail(DTO_B200_ERR_CUDA, "ncclAllGather: %s", a->GetErrorString(e));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    count_launches(ctx, 1);  // NCCL's own kernel; counted so that callers can tell it ran
    return DTO_B200_OK;
}

}  // extern "C"
