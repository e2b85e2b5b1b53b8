/*
 * dto_oracle.h -- CPU ORACLE for the Dual Threshold Optimization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (dual_threshold_optimization_b200/,
 * include/, the CUDA library) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / the timed CPU baseline.
 *
 * It is a plain-C restatement of the reference's algorithm (BrentLab/Dual_Threshold_Optimization
 * crate v2.0.1, pure Rust; no Rust toolchain exists in this image so the reference itself cannot be
 * built -- see DESIGN.md).  The p-value arithmetic lives in the un-vendored dependency
 * statrs 0.17.1 (Cargo.lock:734-736); its published algorithm is restated here and PINNED against
 * every golden the reference's own tests hold for this path (tests/test_oracle_goldens.py).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef DTO_ORACLE_H
#define DTO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors dto/results_objects.rs:22-32 (OptimizationResultRecord, feature_sets omitted). */
typedef struct {
    uint32_t rank1;
    uint32_t rank2;
    uint32_t set1_len;
    uint32_t set2_len;
    uint32_t intersection_size;
    uint32_t permuted;
    uint64_t population_size;
    double pvalue;
} oracle_record_t;

/* ---- statrs 0.17.1 restatement (function::gamma, function::factorial, distribution::Hypergeometric) */
double oracle_ln_gamma(double x);
double oracle_ln_factorial(uint64_t x);
double oracle_ln_binomial(uint64_t n, uint64_t k);
/* Hypergeometric::new(N,K,n).sf(x); returns NaN when new() would fail (K>N or n>N). */
double oracle_hypergeom_sf(uint64_t N, uint64_t K, uint64_t n, uint64_t x);
/* stat_operations/hypergeometric_pvalue.rs:33-50 */
double oracle_hypergeometric_pvalue(uint64_t N, uint64_t K, uint64_t n, uint64_t k);
/* Same value as oracle_hypergeometric_pvalue but with a cached ln_factorial table and the
 * bit-preserving early exit (terms below half an ulp of the accumulator cannot change it).
 * lf must hold ln_factorial(0..N).  Used by the integer-mode grid. */
double oracle_hypergeometric_pvalue_cached(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k);
/* natural log of the same upper tail, computed by log-sum-exp (finite where the p-value underflows) */
double oracle_hypergeometric_log_pvalue(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k);
/* number of tail terms needed to converge to 2^-53 relative (SURVEY 8(d) "R"); 0 for short-circuit cells */
uint64_t oracle_tail_terms(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k);
void oracle_fill_ln_factorial(double *lf, uint64_t N);

/* ---- collections/ranked.rs */
/* sort_genes_and_ranks (ranked.rs:527-542): stable sort by rank; order_out[j] = original index at sorted position j */
void oracle_stable_sort_by_rank(const uint32_t *ranks, size_t n, uint32_t *sorted_ranks, uint32_t *order_out);
/* generate_thresholds (ranked.rs:359-375, INCLUDING the no-op "last = max_rank" bug). Returns count; writes up to cap. */
size_t oracle_generate_thresholds(const uint32_t *sorted_ranks, size_t n, uint32_t *out, size_t cap);

/* ---- reference-faithful path (string ids, per-cell hash set, uncached ln_gamma, full tail) ----
 * process_threshold_pairs (dto/process_threshold_pairs.rs:71-131).  Lists are given already sorted
 * by rank (ids[j], ranks[j] at sorted position j).  perm1/perm2 (may be NULL = unpermuted view)
 * are the `indices` of PermutedRankedFeatureList (collections/permuted.rs:56-60,90-101): position j keeps
 * ranks[j] and holds gene ids[perm[j]].  records_out must hold T1*T2 records, row-major (t1 outer).
 * Returns 0, or -1 if Hypergeometric::new would panic. */
int oracle_process_threshold_pairs_faithful(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                            const uint32_t *thr1, size_t T1,
                                            const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                            const uint32_t *thr2, size_t T2,
                                            const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                                            uint64_t population, oracle_record_t *records_out);

/* rows i % row_stride == 0 only: bounded sample for timing (see dto_oracle.c) */
int oracle_process_threshold_pairs_faithful_sampled(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                                    const uint32_t *thr1, size_t T1,
                                                    const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                                    const uint32_t *thr2, size_t T2,
                                                    const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                                                    uint64_t population, size_t row_stride, oracle_record_t *records_out);

/* optimize's reduction (dto/optimize_main.rs:73-116) over a row-major record vector. Returns index of the winner. */
size_t oracle_argmin_tiebreak(const oracle_record_t *records, size_t count);

/* ---- integer-id path (same results, fast enough for N=6k..60k checking) ----
 * slot2_of_1[a] = list-2 sorted slot of the gene at list-1 sorted slot a, or -1 if absent from list 2.
 * Duplicate ids are not supported in this mode.
 * overlap_out (T1*T2 u32), p_out (T1*T2 f64), logp_out (T1*T2 f64) may each be NULL. best_out may be NULL. */
int oracle_grid_int(const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                    const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                    const int32_t *slot2_of_1,
                    const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                    uint64_t population, const double *lf,
                    uint32_t *overlap_out, double *p_out, double *logp_out, oracle_record_t *best_out);

/* P permuted tasks with caller-supplied indices (perm1: P x n1, perm2: P x n2), static chunks on num_threads OS threads
 * (run/single_node.rs:94-133); results_out[t] = Best record of task t.  For the deep GPU parity tests. */
int oracle_grid_int_batch(const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                          const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                          const int32_t *slot2_of_1, const uint32_t *perm1, const uint32_t *perm2, size_t P,
                          uint64_t population, const double *lf, size_t num_threads, oracle_record_t *results_out);

/* ---- epilogue: stat_operations/fdr.rs:29-60, stat_operations/empirical_pvalue.rs:109-187 */
double oracle_fdr(uint64_t list1_len, uint64_t list2_len, uint64_t overlap, uint64_t population, double sensitivity);
double oracle_empirical_pvalue(const double *permuted_minp, size_t P, double unpermuted_p);

/* ---- CPU timing baseline: run/single_node.rs:83-137 (static ceil(tasks/threads) chunks on threads),
 * each task = faithful process_threshold_pairs + optimize with a fresh Fisher-Yates permutation pair
 * (collections/permuted.rs:56-60; RNG = splitmix/xoshiro seeded per task since thread_rng is unseedable).
 * task_permute[t] != 0 -> permuted.  results_out[t] is the Best record of task t.  mode 0 = faithful, 1 = integer. */
int oracle_run_single_node(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                           const char *const *ids2, const uint32_t *ranks2, size_t n2,
                           const int32_t *slot2_of_1, uint64_t population,
                           const uint8_t *task_permute, size_t n_tasks, size_t num_threads,
                           uint64_t seed, int mode, oracle_record_t *results_out);

int oracle_run_single_node_sampled(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                   const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                   const int32_t *slot2_of_1, uint64_t population,
                                   const uint8_t *task_permute, size_t n_tasks, size_t num_threads,
                                   uint64_t seed, int mode, size_t row_stride, oracle_record_t *results_out);

void oracle_grid_tail_terms(const uint32_t *overlap, const uint32_t *c1, size_t T1, const uint32_t *c2, size_t T2,
                            uint64_t population, const double *lf, uint64_t *terms_out, uint64_t *evaluated_cells_out);

/* uniform Fisher-Yates in rand 0.8.5's loop order (for i in (1..n).rev() swap(i, gen_range(0..=i))) */
void oracle_shuffle(uint32_t *idx, size_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
