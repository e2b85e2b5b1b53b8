/*
 * dto_oracle.c -- CPU ORACLE (test infrastructure, see dto_oracle.h).  NOT product code.
 *
 * Plain-C restatement of the reference hot path.  Build with -ffp-contract=off: the reference
 * (rustc) never fuses a*b+c, and parity of the last bit of ln_gamma depends on it.
 * libm log/exp are the ones Rust's f64::ln / f64::exp bind to on Linux (glibc).
 */
#include "dto_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * statrs 0.17.1  function::gamma::ln_gamma  (Lanczos, g = 10.900511, 11 coefficients)
 * call chain in the reference: stat_operations/hypergeometric_pvalue.rs:40,49 -> Hypergeometric::sf
 *   -> factorial::ln_binomial -> factorial::ln_factorial -> gamma::ln_gamma
 * ---------------------------------------------------------------------------------------------- */
static const double GAMMA_R = 10.900511;
static const double GAMMA_DK[11] = {
    2.48574089138753565546e-5,  1.05142378581721974210,     -3.45687097222016235469,
    4.51227709466894823700,     -2.98285225323576655721,    1.05639711577126713077,
    -1.95428773191645869583e-1, 1.70970543404441224307e-2,  -5.71926117404305781283e-4,
    4.63399473359905636708e-6,  -2.71994908488607703910e-9,
};
static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;
static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;
#define ORACLE_E 2.71828182845904523536028747135266250
#define ORACLE_PI 3.14159265358979323846264338327950288

double oracle_ln_gamma(double x) {
    if (x < 0.5) {
        double s = GAMMA_DK[0];
        for (int i = 1; i <= 10; ++i) s += GAMMA_DK[i] / ((double)i - x);
        return LN_PI - log(sin(ORACLE_PI * x)) - log(s) - LN_2_SQRT_E_OVER_PI -
               (0.5 - x) * log((0.5 - x + GAMMA_R) / ORACLE_E);
    }
    double s = GAMMA_DK[0];
    for (int i = 1; i <= 10; ++i) s += GAMMA_DK[i] / (x + (double)i - 1.0);
    return log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * log((x - 0.5 + GAMMA_R) / ORACLE_E);
}

/* statrs function::factorial: FCACHE[0]=1, FCACHE[i]=FCACHE[i-1]*i for i<=170; ln_factorial(x<=170)=ln(FCACHE[x]) */
static double FCACHE[171];
static pthread_once_t fcache_once = PTHREAD_ONCE_INIT;
static void fcache_init(void) {
    FCACHE[0] = 1.0;
    for (int i = 1; i <= 170; ++i) FCACHE[i] = FCACHE[i - 1] * (double)i;
}

double oracle_ln_factorial(uint64_t x) {
    pthread_once(&fcache_once, fcache_init);
    if (x <= 170) return log(FCACHE[x]);
    return oracle_ln_gamma((double)x + 1.0);
}

double oracle_ln_binomial(uint64_t n, uint64_t k) {
    if (k > n) return -INFINITY;
    return oracle_ln_factorial(n) - oracle_ln_factorial(k) - oracle_ln_factorial(n - k);
}

/* statrs distribution::Hypergeometric::sf: direct ascending upper-tail sum (NOT 1-cdf) */
double oracle_hypergeom_sf(uint64_t N, uint64_t K, uint64_t n, uint64_t x) {
    if (K > N || n > N) return NAN; /* Hypergeometric::new -> Err -> expect() panic in the reference */
    uint64_t mn = (n + K > N) ? (n + K - N) : 0; /* (draws + successes).saturating_sub(population) */
    uint64_t mx = K < n ? K : n;
    if (x < mn) return 1.0;
    if (x >= mx) return 0.0;
    double ln_denom = oracle_ln_binomial(N, n);
    double acc = 0.0;
    for (uint64_t i = x + 1; i <= mx; ++i)
        acc += exp(oracle_ln_binomial(K, i) + oracle_ln_binomial(N - K, n - i) - ln_denom);
    return acc;
}

/* stat_operations/hypergeometric_pvalue.rs:33-50 */
double oracle_hypergeometric_pvalue(uint64_t N, uint64_t K, uint64_t n, uint64_t k) {
    if (K > N || n > N) return NAN;
    if (k == 0) return 1.0;
    return oracle_hypergeom_sf(N, K, n, k - 1);
}

void oracle_fill_ln_factorial(double *lf, uint64_t N) {
    for (uint64_t x = 0; x <= N; ++x) lf[x] = oracle_ln_factorial(x);
}

static inline double lnb_cached(const double *lf, uint64_t a, uint64_t b) {
    if (b > a) return -INFINITY;
    return lf[a] - lf[b] - lf[a - b];
}

/* Same arithmetic, table-cached lf (bit-identical values) + bit-preserving early exit:
 * past the mode the terms fall monotonically; once one is below 2^-55 of the accumulator every
 * remaining `acc += term` rounds back to acc, so stopping changes no bit of the result. */
double oracle_hypergeometric_pvalue_cached(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k) {
    if (K > N || n > N) return NAN;
    if (k == 0) return 1.0;
    uint64_t x = k - 1;
    uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    uint64_t mx = K < n ? K : n;
    if (x < mn) return 1.0;
    if (x >= mx) return 0.0;
    double ln_denom = lnb_cached(lf, N, n);
    uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    double acc = 0.0;
    for (uint64_t i = x + 1; i <= mx; ++i) {
        double term = exp(lnb_cached(lf, K, i) + lnb_cached(lf, N - K, n - i) - ln_denom);
        acc += term;
        if (i > mode + 1 && term <= acc * 0x1p-55) break;
    }
    return acc;
}

double oracle_hypergeometric_log_pvalue(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k) {
    if (K > N || n > N) return NAN;
    if (k == 0) return 0.0;
    uint64_t x = k - 1;
    uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    uint64_t mx = K < n ? K : n;
    if (x < mn) return 0.0;
    if (x >= mx) return -INFINITY;
    double ln_denom = lnb_cached(lf, N, n);
    uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    double M = -INFINITY, S = 0.0; /* sum = exp(M) * S */
    for (uint64_t i = x + 1; i <= mx; ++i) {
        double a = lnb_cached(lf, K, i) + lnb_cached(lf, N - K, n - i) - ln_denom;
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

uint64_t oracle_tail_terms(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k) {
    if (K > N || n > N || k == 0) return 0;
    uint64_t x = k - 1;
    uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    uint64_t mx = K < n ? K : n;
    if (x < mn || x >= mx) return 0;
    double ln_denom = lnb_cached(lf, N, n);
    uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    double M = -INFINITY, S = 0.0;
    uint64_t r = 0;
    for (uint64_t i = x + 1; i <= mx; ++i) {
        double a = lnb_cached(lf, K, i) + lnb_cached(lf, N - K, n - i) - ln_denom;
        ++r;
        if (a > M) {
            S = S * exp(M - a) + 1.0;
            M = a;
        } else {
            S += exp(a - M);
        }
        if (i > mode + 1 && exp(a - M) < S * 0x1p-53) break;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * collections/ranked.rs
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t rank;
    uint32_t idx;
} rank_idx_t;

static int cmp_rank_idx(const void *a, const void *b) {
    const rank_idx_t *x = (const rank_idx_t *)a, *y = (const rank_idx_t *)b;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1; /* == stability of sort_by_key (ranked.rs:534) */
    return 0;
}

void oracle_stable_sort_by_rank(const uint32_t *ranks, size_t n, uint32_t *sorted_ranks, uint32_t *order_out) {
    rank_idx_t *v = (rank_idx_t *)malloc((n ? n : 1) * sizeof(rank_idx_t));
    for (size_t i = 0; i < n; ++i) {
        v[i].rank = ranks[i];
        v[i].idx = (uint32_t)i;
    }
    qsort(v, n, sizeof(rank_idx_t), cmp_rank_idx);
    for (size_t i = 0; i < n; ++i) {
        sorted_ranks[i] = v[i].rank;
        order_out[i] = v[i].idx;
    }
    free(v);
}

/* ranked.rs:359-375.  `current` is u32; `(current as f64 * 1.01 + 1.0).floor() as u32` saturates.
 * The "set the final threshold to the maximum rank" block (:370-372) mutates the OLD self.thresholds
 * (empty at construction), so it is a no-op: the last threshold is generally < max_rank. Reproduced. */
size_t oracle_generate_thresholds(const uint32_t *sorted_ranks, size_t n, uint32_t *out, size_t cap) {
    uint32_t max_rank = n ? sorted_ranks[n - 1] : 0;
    uint32_t current = 1;
    size_t count = 0;
    while (current <= max_rank) {
        if (out && count < cap) out[count] = current;
        ++count;
        double next = floor((double)current * 1.01 + 1.0);
        if (next >= 4294967295.0) break; /* `as u32` saturates at u32::MAX; stop instead of spinning */
        current = (uint32_t)next;
    }
    return count;
}

/* ------------------------------------------------------------------------------------------------
 * reference-faithful grid: string features, cloned per threshold, HashSet<&str> rebuilt per cell
 * (stat_operations/intersect_genes.rs:38-56; Rust's default hasher is SipHash-1-3, restated below)
 * ---------------------------------------------------------------------------------------------- */
#define ROTL64(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND            \
    do {                    \
        v0 += v1;           \
        v1 = ROTL64(v1, 13); \
        v1 ^= v0;           \
        v0 = ROTL64(v0, 32); \
        v2 += v3;           \
        v3 = ROTL64(v3, 16); \
        v3 ^= v2;           \
        v0 += v3;           \
        v3 = ROTL64(v3, 21); \
        v3 ^= v0;           \
        v2 += v1;           \
        v1 = ROTL64(v1, 17); \
        v1 ^= v2;           \
        v2 = ROTL64(v2, 32); \
    } while (0)

static uint64_t siphash13(const unsigned char *in, size_t len, uint64_t k0, uint64_t k1) {
    uint64_t v0 = 0x736f6d6570736575ULL ^ k0, v1 = 0x646f72616e646f6dULL ^ k1;
    uint64_t v2 = 0x6c7967656e657261ULL ^ k0, v3 = 0x7465646279746573ULL ^ k1;
    const unsigned char *end = in + len - (len % 8);
    uint64_t b = ((uint64_t)len) << 56;
    for (; in != end; in += 8) {
        uint64_t m;
        memcpy(&m, in, 8);
        v3 ^= m;
        SIPROUND;
        v0 ^= m;
    }
    uint64_t t = 0;
    memcpy(&t, in, len % 8);
    b |= t;
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

typedef struct {
    const char **slots;
    size_t cap; /* power of two */
} strset_t;

static void strset_build(strset_t *s, char *const *items, size_t n) {
    size_t cap = 8;
    while (cap < 2 * n + 2) cap <<= 1;
    s->cap = cap;
    s->slots = (const char **)calloc(cap, sizeof(char *));
    for (size_t i = 0; i < n; ++i) {
        size_t len = strlen(items[i]);
        size_t h = (size_t)siphash13((const unsigned char *)items[i], len, 0x0706050403020100ULL, 0x0f0e0d0c0b0a0908ULL) & (cap - 1);
        for (;;) {
            if (!s->slots[h]) {
                s->slots[h] = items[i];
                break;
            }
            if (strcmp(s->slots[h], items[i]) == 0) break; /* set semantics: duplicates collapse */
            h = (h + 1) & (cap - 1);
        }
    }
}

static int strset_contains(const strset_t *s, const char *key) {
    size_t len = strlen(key);
    size_t h = (size_t)siphash13((const unsigned char *)key, len, 0x0706050403020100ULL, 0x0f0e0d0c0b0a0908ULL) & (s->cap - 1);
    for (;;) {
        if (!s->slots[h]) return 0;
        if (strcmp(s->slots[h], key) == 0) return 1;
        h = (h + 1) & (s->cap - 1);
    }
}

typedef struct {
    char **items; /* cloned Strings, like Vec<Feature> */
    size_t len;
} featvec_t;

/* get_feature_set_by_threshold: ranked.rs:561-568 (unpermuted), permuted.rs:90-101 (permuted view) */
static featvec_t feature_set_by_threshold(const char *const *ids, const uint32_t *ranks, size_t n,
                                          const uint32_t *perm, uint32_t threshold) {
    featvec_t v;
    v.items = (char **)malloc((n ? n : 1) * sizeof(char *));
    v.len = 0;
    for (size_t j = 0; j < n; ++j) {
        if (ranks[j] <= threshold) {
            const char *src = ids[perm ? perm[j] : j];
            size_t len = strlen(src) + 1;
            char *c = (char *)malloc(len);
            memcpy(c, src, len);
            v.items[v.len++] = c;
        }
    }
    return v;
}

static void featvec_free(featvec_t *v) {
    for (size_t i = 0; i < v->len; ++i) free(v->items[i]);
    free(v->items);
    v->items = NULL;
    v->len = 0;
}

/* intersect_genes.rs:38-56: HashSet of list-1 ids, count list-2 items that hit (list-2 duplicates count multiply) */
static size_t intersect_genes(const featvec_t *g1, const featvec_t *g2) {
    strset_t s;
    strset_build(&s, g1->items, g1->len);
    size_t k = 0;
    for (size_t i = 0; i < g2->len; ++i) k += (size_t)strset_contains(&s, g2->items[i]);
    free((void *)s.slots);
    return k;
}

int oracle_process_threshold_pairs_faithful(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                            const uint32_t *thr1, size_t T1,
                                            const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                            const uint32_t *thr2, size_t T2,
                                            const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                                            uint64_t population, oracle_record_t *out) {
    return oracle_process_threshold_pairs_faithful_sampled(ids1, ranks1, n1, thr1, T1, ids2, ranks2, n2, thr2, T2, perm1,
                                                           perm2, permuted_flag, population, 1, out);
}

/* Same loop restricted to rows i % row_stride == 0 (records of skipped rows are left untouched): a bounded,
 * evenly spread SAMPLE of the reference's work, used only to time the CPU baseline at sizes where one full
 * task costs minutes (bench.py says so in cpu_baseline.sample). row_stride == 1 is the full reference loop. */
int oracle_process_threshold_pairs_faithful_sampled(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                                    const uint32_t *thr1, size_t T1,
                                                    const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                                    const uint32_t *thr2, size_t T2,
                                                    const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                                                    uint64_t population, size_t row_stride, oracle_record_t *out) {
    if (row_stride == 0) row_stride = 1;
    featvec_t *cache = (featvec_t *)calloc(T2 ? T2 : 1, sizeof(featvec_t));
    unsigned char *have = (unsigned char *)calloc(T2 ? T2 : 1, 1);
    int rc = 0;
    for (size_t i = 0; i < T1 && rc == 0; i += row_stride) {
        featvec_t g1 = feature_set_by_threshold(ids1, ranks1, n1, perm1, thr1[i]);
        for (size_t j = 0; j < T2; ++j) {
            if (!have[j]) { /* feature_sets_cache.entry(threshold2).or_insert_with (process_threshold_pairs.rs:92-98) */
                cache[j] = feature_set_by_threshold(ids2, ranks2, n2, perm2, thr2[j]);
                have[j] = 1;
            }
            size_t k = intersect_genes(&g1, &cache[j]);
            double p = oracle_hypergeometric_pvalue(population, g1.len, cache[j].len, k);
            if (isnan(p)) {
                rc = -1;
                break;
            }
            oracle_record_t *r = &out[i * T2 + j];
            r->rank1 = thr1[i];
            r->rank2 = thr2[j];
            r->set1_len = (uint32_t)g1.len;
            r->set2_len = (uint32_t)cache[j].len;
            r->population_size = population;
            r->intersection_size = (uint32_t)k;
            r->pvalue = p;
            r->permuted = permuted_flag ? 1u : 0u;
        }
        featvec_free(&g1);
    }
    for (size_t j = 0; j < T2; ++j)
        if (have[j]) featvec_free(&cache[j]);
    free(cache);
    free(have);
    return rc;
}

/* dto/optimize_main.rs:73-116: min by f64::min fold, keep pvalue == min (exact), keep max intersection,
 * stable sort by (rank1, rank2), first.  Row-major input order == ascending (rank1, rank2) already, so the
 * stable sort's winner is the first surviving record with the lexicographically smallest key. */
size_t oracle_argmin_tiebreak(const oracle_record_t *rec, size_t count) {
    double min_p = INFINITY;
    for (size_t i = 0; i < count; ++i) min_p = fmin(min_p, rec[i].pvalue);
    uint32_t max_k = 0;
    size_t n_min = 0;
    for (size_t i = 0; i < count; ++i)
        if (rec[i].pvalue == min_p) {
            ++n_min;
            if (rec[i].intersection_size > max_k) max_k = rec[i].intersection_size;
        }
    size_t best = (size_t)-1;
    for (size_t i = 0; i < count; ++i) {
        if (rec[i].pvalue != min_p) continue;
        if (n_min > 1 && rec[i].intersection_size != max_k) continue;
        if (best == (size_t)-1 || rec[i].rank1 < rec[best].rank1 ||
            (rec[i].rank1 == rec[best].rank1 && rec[i].rank2 < rec[best].rank2))
            best = i;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * integer-id grid (SURVEY App. B "equivalent integer form"): histogram over (bin1, bin2) of the common
 * genes, 2-D inclusive prefix sum, set sizes = #{ranks <= t}.  Cross-checked against the faithful path in
 * tests/test_oracle_goldens.py.
 * ---------------------------------------------------------------------------------------------- */
static int32_t bin_of_rank(const uint32_t *thr, size_t T, uint32_t rank) {
    size_t lo = 0, hi = T; /* first i with thr[i] >= rank */
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (thr[mid] >= rank) hi = mid;
        else lo = mid + 1;
    }
    return lo < T ? (int32_t)lo : -1;
}

int oracle_grid_int(const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                    const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                    const int32_t *slot2_of_1,
                    const uint32_t *perm1, const uint32_t *perm2, int permuted_flag,
                    uint64_t population, const double *lf,
                    uint32_t *overlap_out, double *p_out, double *logp_out, oracle_record_t *best_out) {
    if (T1 == 0 || T2 == 0) return -2;
    uint32_t *c1 = (uint32_t *)calloc(T1, sizeof(uint32_t));
    uint32_t *c2 = (uint32_t *)calloc(T2, sizeof(uint32_t));
    for (size_t i = 0; i < T1; ++i) {
        uint32_t c = 0;
        for (size_t j = 0; j < n1; ++j) c += ranks1[j] <= thr1[i];
        c1[i] = c;
    }
    for (size_t i = 0; i < T2; ++i) {
        uint32_t c = 0;
        for (size_t j = 0; j < n2; ++j) c += ranks2[j] <= thr2[i];
        c2[i] = c;
    }
    /* position of every list-2 slot in the (permuted) list 2: slot perm2[j'] sits at position j' */
    uint32_t *pos2_of_slot = (uint32_t *)malloc((n2 ? n2 : 1) * sizeof(uint32_t));
    for (size_t j = 0; j < n2; ++j) pos2_of_slot[perm2 ? perm2[j] : j] = (uint32_t)j;
    uint32_t *H = (uint32_t *)calloc(T1 * T2, sizeof(uint32_t));
    for (size_t j = 0; j < n1; ++j) {
        size_t slot1 = perm1 ? perm1[j] : j; /* gene at list-1 position j */
        int32_t slot2 = slot2_of_1[slot1];
        if (slot2 < 0) continue;
        int32_t b1 = bin_of_rank(thr1, T1, ranks1[j]);
        int32_t b2 = bin_of_rank(thr2, T2, ranks2[pos2_of_slot[slot2]]);
        if (b1 < 0 || b2 < 0) continue;
        H[(size_t)b1 * T2 + (size_t)b2] += 1;
    }
    for (size_t i = 0; i < T1; ++i) {
        uint32_t run = 0;
        for (size_t j = 0; j < T2; ++j) {
            run += H[i * T2 + j];
            H[i * T2 + j] = run + (i ? H[(i - 1) * T2 + j] : 0);
        }
    }
    int rc = 0;
    oracle_record_t best;
    memset(&best, 0, sizeof(best));
    int have_best = 0;
    for (size_t i = 0; i < T1 && rc == 0; ++i) {
        for (size_t j = 0; j < T2; ++j) {
            uint32_t k = H[i * T2 + j];
            if (overlap_out) overlap_out[i * T2 + j] = k;
            if (c1[i] > population || c2[j] > population) {
                rc = -1;
                break;
            }
            double p = oracle_hypergeometric_pvalue_cached(lf, population, c1[i], c2[j], k);
            if (p_out) p_out[i * T2 + j] = p;
            if (logp_out) logp_out[i * T2 + j] = oracle_hypergeometric_log_pvalue(lf, population, c1[i], c2[j], k);
            /* row-major scan == ascending (rank1, rank2): strict improvements only keep the first of equals */
            int better = !have_best || p < best.pvalue || (p == best.pvalue && k > best.intersection_size);
            if (better) {
                best.rank1 = thr1[i];
                best.rank2 = thr2[j];
                best.set1_len = c1[i];
                best.set2_len = c2[j];
                best.intersection_size = k;
                best.pvalue = p;
                best.population_size = population;
                best.permuted = permuted_flag ? 1u : 0u;
                have_best = 1;
            }
        }
    }
    if (best_out && rc == 0) *best_out = best;
    free(H);
    free(pos2_of_slot);
    free(c1);
    free(c2);
    return rc;
}

/* SURVEY 8(d) accounting: sum over cells of R (tail terms to converge to 2^-53) and the number of cells the
 * reference does not short-circuit, for a given overlap grid. */
void oracle_grid_tail_terms(const uint32_t *overlap, const uint32_t *c1, size_t T1, const uint32_t *c2, size_t T2,
                            uint64_t population, const double *lf, uint64_t *terms_out, uint64_t *evaluated_cells_out) {
    uint64_t terms = 0, cells = 0;
    for (size_t i = 0; i < T1; ++i)
        for (size_t j = 0; j < T2; ++j) {
            uint64_t r = oracle_tail_terms(lf, population, c1[i], c2[j], overlap[i * T2 + j]);
            terms += r;
            cells += r > 0;
        }
    *terms_out = terms;
    *evaluated_cells_out = cells;
}

/* ------------------------------------------------------------------------------------------------
 * epilogue
 * ---------------------------------------------------------------------------------------------- */
/* stat_operations/fdr.rs:29-60; returns NaN where the reference panics (sensitivity <= 0) */
double oracle_fdr(uint64_t list1_len, uint64_t list2_len, uint64_t overlap, uint64_t population, double sensitivity) {
    if (sensitivity <= 0.0) return NAN;
    double df = fmax((double)overlap / sensitivity, 0.0);
    double b = fmax((double)list1_len - df, 0.0);
    double r = fmax((double)list2_len - df, 0.0);
    double num = b * r;
    double den = (double)population * (double)overlap;
    return den > 0.0 ? num / den : 0.0;
}

/* stat_operations/empirical_pvalue.rs:145-175: #{p_perm <= p_unperm}/P, no +1; P == 0 -> 1.0 */
double oracle_empirical_pvalue(const double *permuted_minp, size_t P, double unpermuted_p) {
    if (P == 0) return 1.0;
    size_t c = 0;
    for (size_t i = 0; i < P; ++i) c += permuted_minp[i] <= unpermuted_p;
    return (double)c / (double)P;
}

/* ------------------------------------------------------------------------------------------------
 * Fisher-Yates in rand 0.8.5 SliceRandom::shuffle order (collections/permuted.rs:58).  thread_rng is
 * OS-seeded ChaCha12 and cannot be seeded, so only the distribution (uniform permutation) is a contract;
 * the generator here is xoshiro256** seeded through splitmix64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t s[4];
} xo_t;
static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void xo_seed(xo_t *r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static uint64_t xo_next(xo_t *r) {
    uint64_t *s = r->s;
    uint64_t result = ROTL64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = ROTL64(s[3], 45);
    return result;
}
/* unbiased integer in [0, bound) */
static uint32_t xo_below(xo_t *r, uint32_t bound) {
    uint64_t m = (uint64_t)(uint32_t)(xo_next(r) >> 32) * (uint64_t)bound;
    uint32_t l = (uint32_t)m;
    if (l < bound) {
        uint32_t t = (uint32_t)(-bound) % bound;
        while (l < t) {
            m = (uint64_t)(uint32_t)(xo_next(r) >> 32) * (uint64_t)bound;
            l = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

static void shuffle_with(uint32_t *idx, size_t n, xo_t *r) {
    for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
    for (size_t i = n; i-- > 1;) {
        uint32_t j = xo_below(r, (uint32_t)i + 1);
        uint32_t t = idx[i];
        idx[i] = idx[j];
        idx[j] = t;
    }
}

void oracle_shuffle(uint32_t *idx, size_t n, uint64_t seed) {
    xo_t r;
    xo_seed(&r, seed);
    shuffle_with(idx, n, &r);
}

/* ------------------------------------------------------------------------------------------------
 * run/single_node.rs:83-137: static chunks of ceil(tasks/threads), one OS thread per chunk, each thread
 * loops optimize(.., debug=false) over its chunk.  (Results there are appended in completion order; here
 * results_out[t] is indexed by task so checks are deterministic -- the consumer only partitions by `permuted`.)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const char *const *ids1;
    const uint32_t *ranks1;
    size_t n1;
    const uint32_t *thr1;
    size_t T1;
    const char *const *ids2;
    const uint32_t *ranks2;
    size_t n2;
    const uint32_t *thr2;
    size_t T2;
    const int32_t *slot2_of_1;
    uint64_t population;
    const double *lf;
    const uint8_t *task_permute;
    size_t begin, end;
    uint64_t seed;
    int mode;
    size_t row_stride;
    oracle_record_t *results;
    int rc;
} chunk_job_t;

static void *chunk_main(void *arg) {
    chunk_job_t *J = (chunk_job_t *)arg;
    uint32_t *perm1 = (uint32_t *)malloc((J->n1 ? J->n1 : 1) * sizeof(uint32_t));
    uint32_t *perm2 = (uint32_t *)malloc((J->n2 ? J->n2 : 1) * sizeof(uint32_t));
    oracle_record_t *grid = NULL;
    if (J->mode == 0) grid = (oracle_record_t *)malloc(J->T1 * J->T2 * sizeof(oracle_record_t));
    for (size_t t = J->begin; t < J->end; ++t) {
        /* process_threshold_pairs.rs:81-82: both permuted views are built on EVERY task, even unpermuted */
        xo_t r;
        xo_seed(&r, J->seed + 0x51ed270b1ULL * (uint64_t)t);
        shuffle_with(perm1, J->n1, &r);
        shuffle_with(perm2, J->n2, &r);
        int permute = J->task_permute[t] != 0;
        if (J->mode == 0) {
            int rc = oracle_process_threshold_pairs_faithful_sampled(J->ids1, J->ranks1, J->n1, J->thr1, J->T1,
                                                                     J->ids2, J->ranks2, J->n2, J->thr2, J->T2,
                                                                     permute ? perm1 : NULL, permute ? perm2 : NULL,
                                                                     permute, J->population, J->row_stride, grid);
            if (rc) {
                J->rc = rc;
                break;
            }
            if (J->row_stride > 1) { /* compact the sampled rows before the reduction */
                size_t rows = 0;
                for (size_t i = 0; i < J->T1; i += J->row_stride, ++rows)
                    if (rows * J->row_stride != rows)
                        memmove(&grid[rows * J->T2], &grid[i * J->T2], J->T2 * sizeof(oracle_record_t));
                J->results[t] = grid[oracle_argmin_tiebreak(grid, rows * J->T2)];
            } else {
                J->results[t] = grid[oracle_argmin_tiebreak(grid, J->T1 * J->T2)];
            }
        } else {
            int rc = oracle_grid_int(J->ranks1, J->n1, J->thr1, J->T1, J->ranks2, J->n2, J->thr2, J->T2,
                                     J->slot2_of_1, permute ? perm1 : NULL, permute ? perm2 : NULL, permute,
                                     J->population, J->lf, NULL, NULL, NULL, &J->results[t]);
            if (rc) {
                J->rc = rc;
                break;
            }
        }
    }
    free(grid);
    free(perm1);
    free(perm2);
    return NULL;
}

int oracle_run_single_node(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                           const char *const *ids2, const uint32_t *ranks2, size_t n2,
                           const int32_t *slot2_of_1, uint64_t population,
                           const uint8_t *task_permute, size_t n_tasks, size_t num_threads,
                           uint64_t seed, int mode, oracle_record_t *results_out) {
    return oracle_run_single_node_sampled(ids1, ranks1, n1, ids2, ranks2, n2, slot2_of_1, population, task_permute,
                                          n_tasks, num_threads, seed, mode, 1, results_out);
}

int oracle_run_single_node_sampled(const char *const *ids1, const uint32_t *ranks1, size_t n1,
                                   const char *const *ids2, const uint32_t *ranks2, size_t n2,
                                   const int32_t *slot2_of_1, uint64_t population,
                                   const uint8_t *task_permute, size_t n_tasks, size_t num_threads,
                                   uint64_t seed, int mode, size_t row_stride, oracle_record_t *results_out) {
    if (n_tasks == 0) return 0;
    if (num_threads == 0) num_threads = 1; /* main.rs:73-76 */
    size_t T1 = oracle_generate_thresholds(ranks1, n1, NULL, 0);
    size_t T2 = oracle_generate_thresholds(ranks2, n2, NULL, 0);
    if (T1 == 0 || T2 == 0) return -2; /* optimize_main.rs:116 unwrap on empty */
    uint32_t *thr1 = (uint32_t *)malloc(T1 * sizeof(uint32_t));
    uint32_t *thr2 = (uint32_t *)malloc(T2 * sizeof(uint32_t));
    oracle_generate_thresholds(ranks1, n1, thr1, T1);
    oracle_generate_thresholds(ranks2, n2, thr2, T2);
    double *lf = NULL;
    if (mode != 0) {
        lf = (double *)malloc((population + 1) * sizeof(double));
        oracle_fill_ln_factorial(lf, population);
    }
    size_t chunk = (n_tasks + num_threads - 1) / num_threads; /* tasks.len().div_ceil(num_threads) */
    size_t n_chunks = (n_tasks + chunk - 1) / chunk;
    chunk_job_t *jobs = (chunk_job_t *)calloc(n_chunks, sizeof(chunk_job_t));
    pthread_t *th = (pthread_t *)calloc(n_chunks, sizeof(pthread_t));
    for (size_t c = 0; c < n_chunks; ++c) {
        chunk_job_t *J = &jobs[c];
        J->ids1 = ids1; J->ranks1 = ranks1; J->n1 = n1; J->thr1 = thr1; J->T1 = T1;
        J->ids2 = ids2; J->ranks2 = ranks2; J->n2 = n2; J->thr2 = thr2; J->T2 = T2;
        J->slot2_of_1 = slot2_of_1; J->population = population; J->lf = lf;
        J->task_permute = task_permute;
        J->begin = c * chunk;
        J->end = (c + 1) * chunk < n_tasks ? (c + 1) * chunk : n_tasks;
        J->seed = seed; J->mode = mode; J->row_stride = row_stride ? row_stride : 1; J->results = results_out; J->rc = 0;
        pthread_create(&th[c], NULL, chunk_main, J);
    }
    int rc = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        pthread_join(th[c], NULL);
        if (jobs[c].rc) rc = jobs[c].rc;
    }
    free(jobs); free(th); free(lf); free(thr1); free(thr2);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Batch form of oracle_grid_int for the deep parity tests: P permuted tasks with caller-supplied indices
 * (perm1: P x n1, perm2: P x n2, row-major; permuted.rs:56-60,90-101 semantics), spread over OS threads the way
 * run/single_node.rs:94-133 spreads tasks.  results_out[t] = the Best record of task t (optimize_main.rs:73-116).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const uint32_t *ranks1, *thr1, *ranks2, *thr2;
    size_t n1, T1, n2, T2;
    const int32_t *slot2_of_1;
    const uint32_t *perm1, *perm2;
    uint64_t population;
    const double *lf;
    size_t begin, end;
    oracle_record_t *results;
    int rc;
} batch_job_t;

static void *batch_main(void *arg) {
    batch_job_t *J = (batch_job_t *)arg;
    for (size_t t = J->begin; t < J->end; ++t) {
        int rc = oracle_grid_int(J->ranks1, J->n1, J->thr1, J->T1, J->ranks2, J->n2, J->thr2, J->T2, J->slot2_of_1,
                                 J->perm1 + t * J->n1, J->perm2 + t * J->n2, 1, J->population, J->lf, NULL, NULL, NULL,
                                 &J->results[t]);
        if (rc) {
            J->rc = rc;
            break;
        }
    }
    return NULL;
}

int oracle_grid_int_batch(const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                          const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2,
                          const int32_t *slot2_of_1, const uint32_t *perm1, const uint32_t *perm2, size_t P,
                          uint64_t population, const double *lf, size_t num_threads, oracle_record_t *results_out) {
    if (P == 0) return 0;
    if (num_threads == 0) num_threads = 1;
    if (num_threads > P) num_threads = P;
    size_t chunk = (P + num_threads - 1) / num_threads;
    size_t n_chunks = (P + chunk - 1) / chunk;
    batch_job_t *jobs = (batch_job_t *)calloc(n_chunks, sizeof(batch_job_t));
    pthread_t *th = (pthread_t *)calloc(n_chunks, sizeof(pthread_t));
    for (size_t c = 0; c < n_chunks; ++c) {
        batch_job_t *J = &jobs[c];
        J->ranks1 = ranks1; J->n1 = n1; J->thr1 = thr1; J->T1 = T1;
        J->ranks2 = ranks2; J->n2 = n2; J->thr2 = thr2; J->T2 = T2;
        J->slot2_of_1 = slot2_of_1; J->perm1 = perm1; J->perm2 = perm2;
        J->population = population; J->lf = lf;
        J->begin = c * chunk;
        J->end = (c + 1) * chunk < P ? (c + 1) * chunk : P;
        J->results = results_out; J->rc = 0;
        pthread_create(&th[c], NULL, batch_main, J);
    }
    int rc = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        pthread_join(th[c], NULL);
        if (jobs[c].rc) rc = jobs[c].rc;
    }
    free(jobs);
    free(th);
    return rc;
}
