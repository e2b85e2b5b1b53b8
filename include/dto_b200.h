/*
 * dto_b200.h -- C ABI of the B200-native Dual Threshold Optimization hot path.
 *
 * Drop-in boundary for the reference crate (BrentLab/Dual_Threshold_Optimization v2.0.1, paths below are
 * relative to the reference root).  The reference has no FFI layer; the two Rust signatures this library
 * sits behind are
 *     dto::optimize(&RankedFeatureList, &RankedFeatureList, permute, population_size, debug)
 *                                                              src/dto/optimize_main.rs:53-59
 *     run::run_single_node(tasks, list1, list2, population_size, num_threads) -> Vec<OptimizationResult>
 *                                                              src/run/single_node.rs:83-89
 * INTEGRATION.md shows the `extern "C"` block a maintainer adds to src/run and src/dto to call these.
 *
 * Conventions: every function returns DTO_B200_OK (0) or a negative error code and never throws across
 * the boundary; dto_b200_last_error() returns a thread-local message for the last failure on the calling
 * thread.  Callers own all host buffers; the library owns all device memory.  A context is bound to ONE
 * CUDA device and may be driven by one host thread at a time (use one context per GPU).
 * There is NO CPU fallback: without a usable CUDA device dto_b200_create fails with DTO_B200_ERR_CUDA.
 */
#ifndef DTO_B200_H
#define DTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTO_B200_OK 0
#define DTO_B200_ERR_INVALID (-1)     /* bad argument (NULL, size mismatch, unsorted ranks, bad permutation ...) */
#define DTO_B200_ERR_CUDA (-2)        /* CUDA runtime failure / no device */
#define DTO_B200_ERR_STATE (-3)       /* call order: no problem set */
#define DTO_B200_ERR_PANIC (-4)       /* a condition on which the reference panics (message mirrors the reference) */
#define DTO_B200_ERR_UNSUPPORTED (-5) /* outside the implemented envelope (dto_b200_get_limits) */
#define DTO_B200_ERR_IO (-6)

/* Mirrors OptimizationResultRecord, src/dto/results_objects.rs:22-32 (feature_sets = FeatureSets::None). */
typedef struct dto_b200_record {
    uint32_t rank1;             /* threshold on list 1 */
    uint32_t rank2;             /* threshold on list 2 */
    uint32_t set1_len;          /* #{ranks1 <= rank1} */
    uint32_t set2_len;          /* #{ranks2 <= rank2} */
    uint32_t intersection_size; /* overlap of the two sets */
    uint32_t flags;             /* DTO_B200_FLAG_* */
    uint64_t population_size;
    double pvalue; /* statrs-order upper-tail hypergeometric p */
} dto_b200_record;

#define DTO_B200_FLAG_PERMUTED 0x1u     /* record.permuted */
/* Exactness of the optimum.  The reference chooses it with `==` / `<` on p-values computed with the HOST libm's exp()
 * (optimize_main.rs:73-80); the device's exp() may differ from it in the last ulp.  The device therefore never decides
 * between two cells whose p-values are closer than 1e-12 relative (+1e-320 absolute) unless their (K, n, k) are equal:
 * such a tie set is re-evaluated on the host in statrs order with the host libm, and the reference's tie-break (largest
 * overlap, then smallest rank1, rank2) is applied to those values.  Every record therefore carries the threshold pair the
 * reference computes on this machine; TIE_RESOLVED merely reports that the host had to settle it. */
#define DTO_B200_FLAG_TIE_RESOLVED 0x2u
#define DTO_B200_FLAG_HOST_PVALUE 0x4u  /* pvalue was evaluated on the host (bit-identical to the reference's on this \
                                           machine); otherwise on the device: same statrs operation order, CUDA exp(), \
                                           within ~1e-14 relative */
#define DTO_B200_FLAG_PATH_FULL 0x8u    /* solved by the full-grid exact pipeline */
/* On records of dto_b200_run_unpermuted (and whatever else went through the full-grid pipeline): the two situations the
 * reference reports on stderr (optimize_main.rs:85-107).  TIE_MINP: several threshold pairs share the minimum p exactly;
 * TIE_OVERLAP: several of them also share the largest overlap, so the smallest (rank1, rank2) was taken.  The CLI prints
 * the reference's notices from these flags for the unpermuted task; permuted tasks do not carry them (the reference
 * prints one notice per tied permutation, thousands of lines on a typical run). */
#define DTO_B200_FLAG_TIE_MINP 0x10u
#define DTO_B200_FLAG_TIE_OVERLAP 0x20u

typedef struct dto_b200_ctx dto_b200_ctx;

/* The implemented envelope (the reference's own types are u32 ranks / usize counts, src/collections/ranked.rs:125-134):
 * inputs beyond it fail with DTO_B200_ERR_UNSUPPORTED, never silently.  List lengths are bounded by the 16-bit packed
 * overlap counters and partner slots of the kernels, thresholds by the per-lane column registers of the scan (a list needs
 * max rank > 7e8 to exceed 2048 thresholds), the population by the host-built ln-factorial table. */
typedef struct dto_b200_limits {
    uint64_t max_features_per_list;   /* 65 534 */
    uint64_t max_thresholds_per_list; /* 2 048 */
    uint64_t max_population;          /* 2^27 */
    uint64_t max_tasks_per_call;
} dto_b200_limits;
int dto_b200_get_limits(dto_b200_limits *out);

const char *dto_b200_last_error(void);
const char *dto_b200_version(void);

/* ---------------------------------------------------------------------------------------------------
 * Low-level engine (integer ids).  Replaces the body of process_threshold_pairs + optimize
 * (src/dto/process_threshold_pairs.rs:71-131, src/dto/optimize_main.rs:60-116) for many tasks per call.
 * ------------------------------------------------------------------------------------------------- */
int dto_b200_device_count(int *count_out);
int dto_b200_create(dto_b200_ctx **ctx_out, int device);
void dto_b200_destroy(dto_b200_ctx *ctx);

/* ranks1/ranks2: ranks in stable-sorted (ascending) order, i.e. RankedFeatureList.ranks after
 *   RankedFeatureList::from (src/collections/ranked.rs:176-191).
 * thr1/thr2: RankedFeatureList.thresholds (src/collections/ranked.rs:359-375), strictly increasing.
 * slot2_of_1[a]: sorted slot in list 2 of the gene held at sorted slot a of list 1, or -1 if list 2 lacks it
 *   (replaces the string HashSet of src/stat_operations/intersect_genes.rs:38-56); must be injective.
 * population: compute_population_size(...) (src/dto/compute_population_size.rs:66-104).
 * Fails with DTO_B200_ERR_PANIC where Hypergeometric::new would panic (a set larger than the population,
 *   src/stat_operations/hypergeometric_pvalue.rs:40-41) or a threshold list is empty (optimize_main.rs:116). */
int dto_b200_set_problem(dto_b200_ctx *ctx, const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                         const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                         const int32_t *slot2_of_1, uint64_t population);

/* optimize(l1, l2, permute=false, N, debug=false) */
int dto_b200_run_unpermuted(dto_b200_ctx *ctx, dto_b200_record *record_out);

/* P tasks of optimize(l1, l2, permute=true, ..) with HOST-SUPPLIED permutation indices: perm1 is P x n1,
 * perm2 is P x n2 (row-major); row p holds the `indices` of PermutedRankedFeatureList
 * (src/collections/permuted.rs:56-60,90-101): sorted position j keeps ranks[j] and holds the gene of slot perm[j].
 * Parity mode: same semantics and the same indices as the oracle. records_out holds P records. */
int dto_b200_run_permuted_indices(dto_b200_ctx *ctx, const uint32_t *perm1, const uint32_t *perm2, size_t P,
                                  dto_b200_record *records_out);

/* P tasks with ON-DEVICE permutations: task p uses permutation id first_perm_id + p, whose pairing is a
 * uniform random permutation obtained by sorting Philox4x32-10 keys (counter = element, stream, id; key = seed;
 * DESIGN.md section 4, K0); results do not depend on batching or on how ids are sharded over GPUs.  Either output may be NULL. HOST buffers. */
int dto_b200_run_permuted_philox(dto_b200_ctx *ctx, uint64_t seed, uint64_t first_perm_id, size_t P,
                                 dto_b200_record *records_out, double *minp_out);

/* Same, but outputs stay in DEVICE memory of ctx's device (d_minp_out: P doubles, d_records_out: P records or
 * NULL) so that a collective (e.g. ncclAllGather issued by the caller) can follow without a host round trip.
 * Synchronous on return. */
int dto_b200_run_permuted_philox_device(dto_b200_ctx *ctx, uint64_t seed, uint64_t first_perm_id, size_t P,
                                        double *d_minp_out, dto_b200_record *d_records_out);

/* The pairing the device generator uses for one permutation id: pos2_of_pos1_out[j] (n1 entries) = sorted
 * position in list 2 paired with sorted position j of list 1, or 0xFFFFFFFF if that position holds a gene
 * absent from list 2.  Lets a checker replay the identical permutation through the oracle. */
int dto_b200_philox_pairing(dto_b200_ctx *ctx, uint64_t seed, uint64_t perm_id, uint32_t *pos2_of_pos1_out);

/* debug=true analogue (optimize_main.rs:68-70): the full T1 x T2 grid, row-major (t1 outer).
 * perm1/perm2 NULL = unpermuted.  Any output may be NULL. logp = natural log of the same tail via log-sum-exp
 * (finite where pvalue underflows to 0). */
int dto_b200_grid_debug(dto_b200_ctx *ctx, const uint32_t *perm1, const uint32_t *perm2, uint32_t *overlap_out,
                        double *pvalue_out, double *logp_out);

/* standalone device evaluation of hypergeometric_pvalue for `count` (N,K,n,k) quadruples
 * (src/stat_operations/hypergeometric_pvalue.rs:33-50); DTO_B200_ERR_PANIC where Hypergeometric::new panics
 * (successes or draws larger than the population, :40-41). */
int dto_b200_hypergeometric_pvalues(dto_b200_ctx *ctx, const uint64_t *N, const uint64_t *K, const uint64_t *n,
                                    const uint64_t *k, size_t count, double *pvalue_out);

/* ---------------------------------------------------------------------------------------------------
 * One process per GPU: the small all-gather of per-permutation minima over NVLink that replaces the MPI gather of
 * src/run/multi_node.rs:148-160.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the host application
 * already mapped, else the system's); DTO_B200_ERR_UNSUPPORTED if there is none.  `nccl_comm` is an ncclComm_t with one
 * rank per GPU: the caller's own, or one made with the two helpers below (unique id = 128 bytes, created on rank 0 and
 * handed to the other ranks by whatever means the application has).
 * allgather_minima: d_send = `count` doubles on ctx's device (e.g. d_minp_out of dto_b200_run_permuted_philox_device),
 * d_recv = count x n_ranks doubles, rank-major.  Enqueued on the library's stream; synchronous on return.
 * ------------------------------------------------------------------------------------------------- */
int dto_b200_nccl_unique_id(void *id_out_128_bytes);
int dto_b200_nccl_comm_create(void **comm_out, int n_ranks, const void *unique_id_128_bytes, int rank, int device);
int dto_b200_nccl_comm_destroy(void *nccl_comm);
int dto_b200_allgather_minima(dto_b200_ctx *ctx, void *nccl_comm, const double *d_send, double *d_recv, size_t count);

typedef struct dto_b200_stats {
    uint64_t tasks_fast;        /* tasks solved by the warp-per-permutation scan kernel */
    uint64_t tasks_full;        /* re-run through the full-grid exact pipeline (min p >= 1 / no candidate) */
    uint64_t candidates;        /* cells evaluated exactly (statrs-order tail) */
    uint64_t level2_cells;      /* cells that passed the critical-overlap screen */
    uint64_t refined_cells;     /* cells whose tail was summed by the ratio recurrence (log p to ~1e-10) */
    uint64_t kernel_launches;   /* launches of this library's kernels since create/reset */
    double last_scan_kernel_ms; /* CUDA-event time of the scan kernel launches of the last run call (sum) */
    double last_sigma_kernel_ms;
    uint64_t last_scan_launches;
    double last_run_ms;         /* CUDA-event time of the whole last run_permuted_* call on the library's stream */
    uint64_t h2d_bytes;         /* bytes copied host->device since create/reset (inputs, tables, task lists) */
    uint64_t d2h_bytes;         /* bytes copied device->host since create/reset (records, status words) */
    uint64_t lptab_entries;     /* size of the per-problem log-p lookup table (8 B each) */
    uint64_t table_cache_hits;  /* set_problem calls that reused the previous problem's screen / log-p tables (same
                                 * population and set sizes per threshold; option "table_cache" = 0 disables) */
    uint64_t tasks_tie_resolved; /* tasks whose optimum was settled on the host (DTO_B200_FLAG_TIE_RESOLVED) */
    uint64_t tie_cells_host;     /* cells re-evaluated on the host for that */
} dto_b200_stats;
int dto_b200_get_stats(dto_b200_ctx *ctx, dto_b200_stats *out);
/* process-wide totals since load, over every context including the pooled ones the host layer uses internally:
 * out3 = {kernel launches of this library, bytes copied host->device, bytes copied device->host} */
int dto_b200_process_totals(uint64_t *out3);
int dto_b200_reset_stats(dto_b200_ctx *ctx);

/* diagnostics: per-task {screened cells, recurrence-refined cells, exactly evaluated cells, then SM cycles / 16 of:
 * the whole task, the histogram scatter, the screen-queue drain, the refine stage, the exact stage} of the LAST scan
 * launch; needs option "task_stats" = 1.  out holds 8 u32 per task. */
int dto_b200_last_batch_task_stats(dto_b200_ctx *ctx, uint32_t *out, size_t max_tasks, size_t *n_out);

/* diagnostics: the per-problem log-p lookup table the scan kernel reads: natural log of
 * hypergeometric_pvalue(N, set1_len[row], set2_len[col], k) for `count` (row, col, k) triples, NaN where (k) lies outside
 * the tabulated range of that cell.  Lets a checker bound the table's error against the oracle. */
int dto_b200_table_logp(dto_b200_ctx *ctx, const uint32_t *row, const uint32_t *col, const uint32_t *k, size_t count,
                        double *logp_out);

/* tunables: "batch" (permutations per launch), "warps_per_cta", "levels" (2..32) and "packed_screen" (before set_problem),
 * "task_stats", "table_cache" (1 = reuse the screen / log-p tables of the previous problem when population and set sizes
 * per threshold are equal, the default; 0 = always rebuild) */
int dto_b200_set_option(dto_b200_ctx *ctx, const char *name, int64_t value);

/* micro-probes used by bench.py for roofline denominators (measured live, not assumed) */
int dto_b200_probe_fp64_tflops(dto_b200_ctx *ctx, double *tflops_out);
int dto_b200_probe_hbm_gbs(dto_b200_ctx *ctx, double *gbs_out);

/* ---------------------------------------------------------------------------------------------------
 * Host layer mirroring the reference's public surface (strings in, records / JSON out).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dto_b200_ranked_list dto_b200_ranked_list;   /* RankedFeatureList, src/collections/ranked.rs:125-134 */
typedef struct dto_b200_feature_list dto_b200_feature_list; /* FeatureList, src/collections/feature_list.rs */

/* RankedFeatureList::from (ranked.rs:176-191): length check, stable sort by rank, thresholds. */
int dto_b200_ranked_list_from(const char *const *ids, const uint32_t *ranks, size_t n, dto_b200_ranked_list **out);
/* read_ranked_feature_list_from_csv (src/read/read_ranked_feature_list_from_csv.rs:50-70) */
int dto_b200_read_ranked_list_csv(const char *path, dto_b200_ranked_list **out);
void dto_b200_ranked_list_free(dto_b200_ranked_list *l);
size_t dto_b200_ranked_list_len(const dto_b200_ranked_list *l);
size_t dto_b200_ranked_list_num_thresholds(const dto_b200_ranked_list *l);
const uint32_t *dto_b200_ranked_list_thresholds(const dto_b200_ranked_list *l);
const uint32_t *dto_b200_ranked_list_ranks(const dto_b200_ranked_list *l);
const char *dto_b200_ranked_list_id(const dto_b200_ranked_list *l, size_t sorted_index);

int dto_b200_feature_list_from(const char *const *ids, size_t n, dto_b200_feature_list **out);
/* read_feature_list_from_file (src/read/read_feature_list_from_file.rs:45-53) */
int dto_b200_read_feature_list(const char *path, dto_b200_feature_list **out);
void dto_b200_feature_list_free(dto_b200_feature_list *l);
size_t dto_b200_feature_list_len(const dto_b200_feature_list *l);
const char *dto_b200_feature_list_id(const dto_b200_feature_list *l, size_t index);

/* compute_population_size (src/dto/compute_population_size.rs:66-104); background may be NULL.
 * The reference's panics come back as DTO_B200_ERR_PANIC with the same message. */
int dto_b200_compute_population_size(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2,
                                     const dto_b200_feature_list *background, uint64_t *population_out);

/* Canonicalise the two lists (string ids -> slot map) and load them into ctx (calls dto_b200_set_problem).
 * Duplicate feature ids inside one list are rejected with DTO_B200_ERR_INVALID (documented deviation:
 * the reference silently mis-counts them). */
int dto_b200_load_lists(dto_b200_ctx *ctx, const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2,
                        uint64_t population);

/* optimize(l1, l2, permute, population, debug=false) for ONE task on `ctx` (optimize_main.rs:53-59). With
 * permute != 0 the permutation is philox id `perm_id` under `seed`. */
int dto_b200_optimize(dto_b200_ctx *ctx, const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, int permute,
                      uint64_t population, uint64_t seed, uint64_t perm_id, dto_b200_record *record_out);

/* run_single_node(tasks, l1, l2, population, num_threads) (src/run/single_node.rs:83-137): task_permute[t] is
 * Task.permute (src/run/task.rs:3-9).  records_out[t] belongs to task t.  Work is sharded over the devices
 * listed (n_devices == 0 -> device 0 only); this replaces both the thread pool and the MPI scatter/gather of
 * src/run/multi_node.rs:114-161 on one box.  Permuted task t uses philox id t under `seed`. */
/* Seeds: the null is a pure function of (seed, Philox id).  The CLI draws `seed` from the OS when --seed is absent (the
 * reference shuffles with thread_rng, permuted.rs:58) and uses ids 1 .. permutations; dto_b200_run_pairs gives pair q the
 * seed  seed + q * 0x9E3779B97F4A7C15  and the same ids, so pair 0 of a batch reproduces the CLI run with that --seed. */
int dto_b200_run_single_node(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, uint64_t population,
                             const uint8_t *task_permute, size_t n_tasks, const int *devices, size_t n_devices,
                             uint64_t seed, dto_b200_record *records_out);

/* The same with explicit Task.id values (src/run/task.rs:3-9): permuted task t uses Philox id task_ids[t] (NULL: id = t).
 * Lets a caller that shards one job over several processes (one rank per GPU, as src/run/multi_node.rs:114-161 shards over
 * MPI ranks) hand each process its id range and still obtain the results of the single-process run. */
int dto_b200_run_tasks(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, uint64_t population,
                       const uint64_t *task_ids, const uint8_t *task_permute, size_t n_tasks, const int *devices,
                       size_t n_devices, uint64_t seed, dto_b200_record *records_out);

/* hypergeometric_pvalue (src/stat_operations/hypergeometric_pvalue.rs:33-50) evaluated on the HOST the way the library's
 * tie resolver and epilogue do: statrs operation order, host-built ln-factorial table, host libm exp().  Bit-identical to
 * the reference on this machine (pinned to the reference's goldens by tests/test_host_layer.py); needs no GPU. */
int dto_b200_hypergeometric_pvalue_host(uint64_t N, uint64_t K, uint64_t n, uint64_t k, double *pvalue_out);

/* fdr (src/stat_operations/fdr.rs:29-60); sensitivity <= 0 -> DTO_B200_ERR_PANIC */
int dto_b200_fdr(uint64_t list1_len, uint64_t list2_len, uint64_t overlap, uint64_t population, double sensitivity,
                 double *fdr_out);

typedef struct dto_b200_final_result { /* the JSON object of empirical_pvalue.rs:145-186 */
    uint64_t rank1, rank2, set1_len, set2_len, population_size, unpermuted_intersection_size;
    double unpermuted_pvalue, empirical_pvalue, fdr;
} dto_b200_final_result;
/* empirical_pvalue (src/stat_operations/empirical_pvalue.rs:109-187).  The count of permuted p <= unpermuted p is the
 * reference's: a permuted record whose p lies within 1e-12 relative of the unpermuted one is compared through host
 * re-evaluations of both (statrs order, host libm), never through device-vs-host last-ulp noise. */
int dto_b200_empirical_pvalue(const dto_b200_record *records, size_t n, dto_b200_final_result *out);
/* serde_json::to_string_pretty of that object (src/main.rs:165): alphabetical keys, 2-space indent, shortest
 * round-trip floats.  Writes at most cap bytes incl. NUL; returns needed length in *len_out. */
int dto_b200_final_result_json(const dto_b200_final_result *r, char *buf, size_t cap, size_t *len_out);

/* Batched list-pair driver (BASELINE config 4: e.g. 2 000 TF binding x perturbation pairs): pair q is the whole CLI run
 * of src/main.rs:91-165 on (lists1[q], lists2[q], populations[q]) -- one unpermuted task + `permutations` permuted tasks
 * + empirical_pvalue -- and yields results_out[q].  Pairs shard over `devices` (contiguous chunks).  Pair q is exactly
 * dto_b200_run_single_node on tasks [unpermuted, permuted x permutations] with seed  seed + q * 0x9E3779B97F4A7C15
 * (permuted task t = 1..permutations uses Philox id t), so pair 0 reproduces the CLI run with the same --seed and no
 * result depends on the device list. */
int dto_b200_run_pairs(const dto_b200_ranked_list *const *lists1, const dto_b200_ranked_list *const *lists2,
                       const uint64_t *populations, size_t n_pairs, size_t permutations, const int *devices,
                       size_t n_devices, uint64_t seed, dto_b200_final_result *results_out);

#ifdef __cplusplus
}
#endif
#endif /* DTO_B200_H */
