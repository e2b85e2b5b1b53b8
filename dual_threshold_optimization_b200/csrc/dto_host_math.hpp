// dto_host_math.hpp -- host-side statrs 0.17.1 ln_factorial (the table the device reads) and small helpers.
//
// The reference evaluates ln_factorial through statrs (function::factorial::ln_factorial ->
// function::gamma::ln_gamma; call site src/stat_operations/hypergeometric_pvalue.rs:40,49).  statrs is not
// vendored in the reference tree (Cargo.lock:734-736), so its published algorithm is restated here:
//   x <= 170 : ln(FCACHE[x]),  FCACHE[0] = 1, FCACHE[i] = FCACHE[i-1] * i
//   else     : ln_gamma(x + 1), Lanczos g = 10.900511 with 11 coefficients, evaluated left to right.
// The table is built ONCE per problem on the host with the libm the reference links (glibc log), so the
// device sums exactly the same ln_factorial values as the reference; only exp() differs (CUDA, <= 1 ulp).
// Compile this translation unit with -ffp-contract=off (rustc never fuses a*b+c).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

namespace dto {

inline double host_ln_gamma(double x) {
    static const double kR = 10.900511;
    static const double kDk[11] = {
        2.48574089138753565546e-5,  1.05142378581721974210,    -3.45687097222016235469,
        4.51227709466894823700,     -2.98285225323576655721,   1.05639711577126713077,
        -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
        4.63399473359905636708e-6,  -2.71994908488607703910e-9,
    };
    static const double kLn2SqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;
    static const double kE = 2.71828182845904523536028747135266250;
    // only x >= 0.5 is reachable from ln_factorial
    double s = kDk[0];
    for (int i = 1; i <= 10; ++i) s += kDk[i] / (x + (double)i - 1.0);
    return std::log(s) + kLn2SqrtEOverPi + (x - 0.5) * std::log((x - 0.5 + kR) / kE);
}

inline void host_fill_ln_factorial(std::vector<double> &lf, uint64_t N) {
    lf.resize(N + 1);
    double f = 1.0;
    for (uint64_t x = 0; x <= N; ++x) {
        if (x <= 170) {
            if (x > 0) f *= (double)x;
            lf[x] = std::log(f);
        } else {
            lf[x] = host_ln_gamma((double)x + 1.0);
        }
    }
}

// hypergeometric_pvalue(N, K, n, k) exactly as the reference evaluates it on THIS host (hypergeometric_pvalue.rs:33-50 ->
// statrs 0.17.1 Hypergeometric::sf): ascending direct sum of exp(lnC(K,i) + lnC(N-K,n-i) - lnC(N,n)) with the host libm's
// exp() -- the function Rust's f64::exp links to.  Same early exit as the device version (bit-preserving: past the mode a
// term below 2^-55 of the accumulator cannot change it, nor can any later one).  `lf` = host_fill_ln_factorial(N).
// Used wherever the reference's comparison of two p-values (== / < in optimize_main.rs:73-80, <= in
// empirical_pvalue.rs:160-165) could hinge on the last ulp of exp(): the device shortlists, the host decides.
inline double host_hypergeom_pvalue_exact(const double *lf, uint64_t N, uint64_t K, uint64_t n, uint64_t k) {
    if (K > N || n > N) return NAN;
    if (k == 0) return 1.0;
    const uint64_t x = k - 1;
    const uint64_t mn = (n + K > N) ? (n + K - N) : 0;
    const uint64_t mx = K < n ? K : n;
    if (x < mn) return 1.0;
    if (x >= mx) return 0.0;
    const double ln_denom = (lf[N] - lf[n]) - lf[N - n];
    const uint64_t mode = (uint64_t)(((double)(n + 1) * (double)(K + 1)) / (double)(N + 2));
    const double lfK = lf[K], lfNK = lf[N - K];
    const uint64_t NK = N - K;
    double acc = 0.0;
    for (uint64_t i = x + 1; i <= mx; ++i) {
        const double a = (lfK - lf[i]) - lf[K - i];
        const uint64_t ni = n - i;
        const double b = (ni > NK) ? -INFINITY : (lfNK - lf[ni]) - lf[NK - ni];
        const double term = std::exp((a + b) - ln_denom);
        acc += term;
        if (i > mode + 1 && term <= acc * 0x1p-55) break;
    }
    return acc;
}

// Two p-values closer than this are "ambiguous": the device's exp() (<= 1 ulp) and the host libm's may order them
// differently (worst case ~1.4e-14 relative over a 60-term tail; the absolute part covers sums of subnormal terms).
constexpr double kTieRel = 1e-12;
constexpr double kTieAbs = 1e-320;
inline bool host_pvalues_ambiguous(double a, double b) {
    const double lo = a < b ? a : b, hi = a < b ? b : a;
    return hi <= lo * (1.0 + kTieRel) + kTieAbs;
}

}  // namespace dto
