// cgg_math.cuh -- per-row GLM log-density terms, prior log-densities and Philox for the CGGibbs
// kernels.  fp64 throughout.  Each function states which reference expression it evaluates.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cggibbs.h"

namespace cgg {

constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;
constexpr double kLogitClampEta = 36.04365338911715;   // -log(DBL_EPSILON)
constexpr double kLogEps = -36.04365338911715;         // log(DBL_EPSILON)

// new_eta = current_eta + X_j * diff (R/glm_utils.R:126-132).  R evaluates the product and the sum
// as two vector ops, i.e. two roundings; an FMA here would make the carried eta drift from the
// reference's, so contraction is blocked explicitly.
__device__ __forceinline__ double eta_shift(double eta, double x, double diff) {
    return __dadd_rn(eta, __dmul_rn(x, diff));
}

// One row's log-density up to a per-dataset constant (added once per sum by ll_finish):
//   gaussian  dnorm(y, eta, sd, log=TRUE)            -> -0.5 z^2            (R/glm_utils.R:40-42)
//   binomial  dbinom(y, 1, logit_linkinv(eta), TRUE) -> -softplus(+-eta)    (R/glm_utils.R:45-47)
//   poisson   dpois(y, pmax(exp(eta), eps), TRUE)    -> y log(mu) - mu      (R/glm_utils.R:50-52)
template <int FAMILY>
__device__ __forceinline__ double row_term(double y, double eta, double inv_sd);

template <>
__device__ __forceinline__ double row_term<CGG_GAUSSIAN>(double y, double eta, double inv_sd) {
    double z = (y - eta) * inv_sd;
    return -0.5 * z * z;
}

// stats' logit_linkinv clamps exp(eta) to [DBL_EPSILON, 1/DBL_EPSILON] when |eta| > 30, which is the
// same as evaluating at eta = -+36.04...; dbinom_raw then returns log(p) (y = 1) or log(1 - p)
// (y = 0).  Both equal -log(1 + exp(-+eta)), evaluated here without forming p or 1 - p:
//   -softplus(s) = -(max(s, 0) + log1p(exp(-|s|))),  s = eta for y = 0, -eta for y = 1.
// NaN propagates (comparisons with NaN are false, exp/log1p keep it).
template <>
__device__ __forceinline__ double row_term<CGG_BINOMIAL>(double y, double eta, double) {
    double s = (y > 0.5) ? -eta : eta;
    s = (s > 30.0) ? kLogitClampEta : ((s < -30.0) ? -kLogitClampEta : s);
    double t = exp(-fabs(s));
    double r = log1p(t);
    return -((s > 0.0 ? s : 0.0) + r);
}

// mu = pmax(exp(eta), eps); dpois_raw(y, mu) = y log(mu) - mu - lgamma(y + 1).  The lgamma term
// does not depend on beta and is added once per sum.  exp overflow gives -Inf as in R
// (!R_FINITE(lambda) -> R_D__0), never NaN.
template <>
__device__ __forceinline__ double row_term<CGG_POISSON>(double y, double eta, double) {
    double le = (eta < kLogEps) ? kLogEps : eta;
    double mu = exp(le);
    double v = y * le - mu;
    return (mu > 1.7976931348623157e308) ? -INFINITY : v;
}

struct PriorParams {
    int kind;
    double mu, sigma, df;
    double c0;       // additive constant of one coordinate's log-density
    double inv_sigma;
};

// distributional::density(beta_prior, x, log = TRUE) for one coordinate (R/glm_utils.R:109)
__device__ __forceinline__ double prior_logdens(const PriorParams &pp, double x) {
    double z = (x - pp.mu) * pp.inv_sigma;
    if (pp.kind == CGG_PRIOR_NORMAL) return pp.c0 - 0.5 * z * z;        // -(ln sqrt(2pi) + z^2/2 + log sigma)
    if (pp.kind == CGG_PRIOR_LAPLACE) return pp.c0 - fabs(z);           // -log(2 sigma) - |x - mu| / sigma
    return pp.c0 - 0.5 * (pp.df + 1.0) * log1p(z * z / pp.df);          // dt(z, df, log=TRUE) - log(sigma)
}

// Philox4x32-10 (Salmon et al. 2011); same definition as oracle.c:orc_philox_uniform.
__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// Uniform #idx of chain `chain` under `seed`, strictly inside (0, 1) like R's runif().
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t chain, uint64_t idx) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = chain, c3 = 0x43474742u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t x = ((uint64_t)c0 << 32) | c1;
    return ((double)(x >> 12) + 0.5) * 0x1p-52;
}

// 128-bit streaming load of two doubles that are never written by any kernel (X columns, y):
// read-only path, do not pollute L1.  `volatile` on purpose: a plain asm counts as side-effect free and
// the compiler then hoists it above the row-range checks (a speculative, possibly misaligned address);
// latency is hidden by the explicit register double-buffering in warp_pass_chain instead.
__device__ __forceinline__ double2 ld_stream2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_ro2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

}  // namespace cgg
