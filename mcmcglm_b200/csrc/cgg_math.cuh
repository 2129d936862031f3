// cgg_math.cuh -- per-row GLM log-density terms, prior log-densities and Philox for the CGGibbs
// kernels.  fp64 throughout.  Each function states which reference expression it evaluates.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cggibbs.h"
#include "cgg_math_tables.cuh"

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

// Fused softplus(s) = max(s, 0) + log1p(exp(-|s|)) for two independent arguments at once (the two rows a
// lane owns in a tile), written 2-wide so that the two dependency chains interleave.  ~40 fp64 operations
// per argument (libdevice exp + log1p cost ~85), <= ~1 ulp:
//   exp(-|s|) : k = rint(-|s| log2 e), r = -|s| - k ln2 (two-step Cody-Waite), degree-11 polynomial, exponent add
//   log1p(t)  : c = rint(32 t)/32, v = (t - c)/(1 + c) via a 33-entry {1/(1+c), log(1+c)} table in shared
//               memory, log1p(t) = log(1+c) + v + v^2 q(v) with |v| <= 1/64
// stats' logit_linkinv clamps exp(eta) to [DBL_EPSILON, 1/DBL_EPSILON] when |eta| > 30, which is the same as
// evaluating at |s| = 36.04...: applied here to a = |s|.  Coefficients and table: cgg_math_tables.cuh
// (tools/gen_math_tables.py, checked against mpmath).  NaN propagates; the table index is clamped so a NaN
// can never index out of bounds.
// exp(r) = pe + r po and exp(-r) = pe - r po: even and odd halves of the degree-11 polynomial (two interleaved
// Horner chains in r^2, depth 7 instead of 11)
__device__ __forceinline__ void poly_exp_eo(double r, double &pe, double &po) {
    static_assert(EXP_DEG == 11, "layout below assumes degree 11");
    const double r2 = r * r;
    pe = EXP_C[10]; po = EXP_C[11];
    pe = fma(pe, r2, EXP_C[8]); po = fma(po, r2, EXP_C[9]);
    pe = fma(pe, r2, EXP_C[6]); po = fma(po, r2, EXP_C[7]);
    pe = fma(pe, r2, EXP_C[4]); po = fma(po, r2, EXP_C[5]);
    pe = fma(pe, r2, EXP_C[2]); po = fma(po, r2, EXP_C[3]);
    pe = fma(pe, r2, EXP_C[0]); po = fma(po, r2, EXP_C[1]);
}
__device__ __forceinline__ double poly_exp(double r) {
    double pe, po;
    poly_exp_eo(r, pe, po);
    return fma(po, r, pe);
}
// v * 2^k by exponent arithmetic (v in [0.5, 2], the result a normal number)
__device__ __forceinline__ double scale2(double v, int k) { return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v)); }

// ---- R's dbinom(0, 1, p, log = TRUE) at large eta ---------------------------------------------------
// For y = 0 R does not evaluate log(1 - p) = -softplus(eta) as a function of eta: stats' logit_linkinv rounds
// p = e / (1 + e), e = exp(eta), to a double first, and nmath's dbinom_raw then takes log(q) of q = 1 - p (exact by
// Sterbenz; reference: R/glm_utils.R:45-47 -> dbinom -> dbinom_raw(x = 0, n = 1, p, q = 1 - p), p >= 0.1 branch).  For
// eta > ~2.2, p lies in [0.9, 1) where doubles are 2^-53 apart, so q carries an absolute error of up to 2^-54 and log(q)
// one of up to 2^-54 / q = 5.6e-17 (1 + e^eta): 1.7e-13 at eta = 8, 5e-8 at eta = 20, 6e-4 at eta = 30.  That is R's
// result, so it is reproduced: with T = exp(-eta), q_hat = T / (1 + T) (1 ulp), p_R = fl(1 - q_hat), q_R = 1 - p_R and
//     log(q_R) = -softplus(eta) + log1p(rho),   rho = (q_R - q_hat) / q_hat = (q_R - q_hat) (1 + T) e^eta,
// |rho| <= 6e-4.  q_R is the very double R forms except when 1 - q lies within ~1e-16 q of a rounding boundary (3 rows
// in 1e4 at eta = 8, none from eta = 15 on; checked against the literal form in tests/test_oracle_math.py).  Applied for
// kRFormLo < eta <= 30; below kRFormLo the two forms differ by < 2e-14 relative per row (covered by the 1e-12 gate), above
// 30 stats' clamp makes p, q constants and both forms give log(2^-52).
constexpr double kRFormLo = 8.0;
__device__ __forceinline__ double rform_log1p_rho(double T /* exp(-eta) < e^-8 */, double Einv /* exp(eta) */) {
    const double u = fma(-T, fma(-T, fma(-T, 1.0 - T, 1.0), 1.0), 1.0);     // 1 / (1 + T), T^5 < 5e-18
    const double qh = T * u;
    const double pR = __dadd_rn(1.0, -qh);
    const double qR = __dadd_rn(1.0, -pR);
    const double rho = (qR - qh) * (Einv * (1.0 + T));
    return rho * fma(-rho, fma(-rho, 1.0 / 3.0, 0.5), 1.0);                  // log1p(rho), rho^4 / 4 < 4e-14 |rho|
}
__device__ __forceinline__ double poly_l1p_q(double v, double v2) {
    static_assert(L1P_QDEG == 6, "layout below assumes degree 6");
    double qe = L1P_Q[6], qo = L1P_Q[5];
    qe = fma(qe, v2, L1P_Q[4]); qo = fma(qo, v2, L1P_Q[3]);
    qe = fma(qe, v2, L1P_Q[2]); qo = fma(qo, v2, L1P_Q[1]);
    qe = fma(qe, v2, L1P_Q[0]);
    return fma(qo, v, qe);
}
// z0 / z1: the row's response is 0 (s = eta): R's log(q) form applies above kRFormLo (see rform_log1p_rho)
// LOGIT (default): stats' logit clamp at |s| > 30.  !LOGIT (negative binomial, log link): plain softplus for any s; for
// |s| > 64 the correction log1p(exp(-|s|)) < 2e-28 is evaluated at 64 (it cannot change the sum).
template <bool LOGIT = true>
__device__ __forceinline__ void softplus2(double s0, double s1, const double2 *tab, double &o0, double &o1, bool z0 = false, bool z1 = false) {
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to nearest integer
    double a0 = fabs(s0), a1 = fabs(s1);
    const double b0 = a0, b1 = a1;              // |s| as it enters the result
    if (LOGIT) { a0 = (a0 > 30.0) ? kLogitClampEta : a0; a1 = (a1 > 30.0) ? kLogitClampEta : a1; }
    else { a0 = (a0 > 64.0) ? 64.0 : a0; a1 = (a1 > 64.0) ? 64.0 : a1; }
    const double kd0 = fma(-a0, 1.4426950408889634, SHIFT), kd1 = fma(-a1, 1.4426950408889634, SHIFT);
    const double kf0 = kd0 - SHIFT, kf1 = kd1 - SHIFT;
    double r0 = fma(kf0, -6.93147180369123816490e-01, -a0), r1 = fma(kf1, -6.93147180369123816490e-01, -a1);
    r0 = fma(kf0, -1.90821492927058770002e-10, r0); r1 = fma(kf1, -1.90821492927058770002e-10, r1);
    double pe0, po0, pe1, po1;
    poly_exp_eo(r0, pe0, po0); poly_exp_eo(r1, pe1, po1);
    const double p0 = fma(po0, r0, pe0), p1 = fma(po1, r1, pe1);
    const double t0 = scale2(p0, __double2loint(kd0)), t1 = scale2(p1, __double2loint(kd1));   // p * 2^k, k in [-52, 0]
    const double md0 = fma(t0, (double)L1P_N, SHIFT), md1 = fma(t1, (double)L1P_N, SHIFT);
    int m0 = __double2loint(md0), m1 = __double2loint(md1);
    m0 = min(max(m0, 0), L1P_N); m1 = min(max(m1, 0), L1P_N);
    const double2 tb0 = tab[m0], tb1 = tab[m1];
    const double v0 = fma(md0 - SHIFT, -1.0 / L1P_N, t0) * tb0.x, v1 = fma(md1 - SHIFT, -1.0 / L1P_N, t1) * tb1.x;
    const double w0 = v0 * v0, w1 = v1 * v1;
    const double q0 = poly_l1p_q(v0, w0), q1 = poly_l1p_q(v1, w1);
    const double l0 = tb0.y + fma(w0, q0, v0), l1 = tb1.y + fma(w1, q1, v1);
    o0 = ((s0 > 0.0) ? (LOGIT ? a0 : b0) : 0.0) + l0;
    o1 = ((s1 > 0.0) ? (LOGIT ? a1 : b1) : 0.0) + l1;
    if (!LOGIT) return;
    const bool f0 = z0 && s0 > kRFormLo && s0 <= 30.0, f1 = z1 && s1 > kRFormLo && s1 <= 30.0;
    if (f0 || f1) {     // rare once a chain is near its stationary region (|eta| of a few units)
        if (f0) o0 -= rform_log1p_rho(t0, scale2(fma(-po0, r0, pe0), -__double2loint(kd0)));
        if (f1) o1 -= rform_log1p_rho(t1, scale2(fma(-po1, r1, pe1), -__double2loint(kd1)));
    }
}

// ---- fp32 "coarse" softplus for the pre-filter ------------------------------------------------------
// softplus32(s) = max(s,0) + log1p(exp(-|s|)) from the hardware approximations ex2.approx / lg2.approx.
// Its absolute error against the exact value is bounded by kCoarseKappa * (1 + |s|) for every fp32 input
// (measured exhaustively over all |s| <= 37 by tests/test_gpu_coarse.py, margin >= 2x), NaN propagates.
// The stats logit clamp (|s| > 30 -> 36.04...) is a discontinuity: a row whose |s| is within 0.02 of 30
// cannot be bounded, `near` reports it and the candidate is then treated as undecided.
constexpr float kCoarseKappa = 4.76837158203125e-07f;   // 2^-21
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float softplus32(float s, bool &near) {
    float a = fabsf(s);
    near = near || (fabsf(a - 30.0f) < 0.02f);
    a = (a > 30.0f) ? 36.0436534f : a;
    const float t = ex2_approx(-a * 1.44269504f);
    const float l = lg2_approx(1.0f + t) * 0.693147181f;
    return ((s > 0.0f) ? a : 0.0f) + l;
}
// The same, for a row with response y = 0 (z): an exact evaluation follows R's log(1 - p) form for 8 < s <= 30
// (rform_log1p_rho), which deviates from the smooth -softplus by at most 2^-54 (1 + e^s); `noise` accumulates an upper
// bound of e^s over those (row, candidate) pairs (margins for the fp32 rounding of s; e^s = 1 / exp(-s), rounded up).
__device__ __forceinline__ float softplus32n(float s, bool z, bool &near, float &noise) {
    float a = fabsf(s);
    near = near || (fabsf(a - 30.0f) < 0.02f);
    const bool rf = z && s > 7.9f && s <= 30.02f;
    a = (a > 30.0f) ? 36.0436534f : a;
    const float t = ex2_approx(-a * 1.44269504f);
    if (rf) noise += __frcp_ru(t) * 1.001f;
    const float l = lg2_approx(1.0f + t) * 0.693147181f;
    return ((s > 0.0f) ? a : 0.0f) + l;
}

// The two rows (i, i+1) a lane owns in a tile, with everything that does not depend on the candidate
// hoisted out of the candidate loop.  term(delta) returns the sum of the two rows' log-density terms up to
// a per-dataset constant (added once per sum, ll_const):
//   gaussian  dnorm(y, eta', sd, log=TRUE)            -> -0.5 z^2            (R/glm_utils.R:40-42)
//   binomial  dbinom(y, 1, logit_linkinv(eta'), TRUE) -> -softplus(+-eta')   (R/glm_utils.R:45-47; y = 0 and
//             eta' > 8: plus the rounding of p that R's log(1 - p) carries, rform_log1p_rho)
//   poisson   dpois(y, pmax(exp(eta'), eps), TRUE)    -> y log(mu) - mu      (R/glm_utils.R:50-52)
// with eta' = eta + X_j * delta formed as two roundings (eta_shift).
template <int FAMILY> struct RowPair;

template <> struct RowPair<CGG_GAUSSIAN> {
    double y0, y1, e0, e1, x0, x1;
    __device__ __forceinline__ RowPair(double2 y, double2 e, double2 x) : y0(y.x), y1(y.y), e0(e.x), e1(e.y), x0(x.x), x1(x.y) {}
    __device__ __forceinline__ double term(double dk, double inv_sd, const double2 *) const {
        const double z0 = (y0 - eta_shift(e0, x0, dk)) * inv_sd, z1 = (y1 - eta_shift(e1, x1, dk)) * inv_sd;
        return -0.5 * fma(z0, z0, z1 * z1);
    }
};

// dbinom_raw returns log(p) (y = 1) or log(1 - p) (y = 0) = -softplus(s), s = eta' for y = 0, -eta' for
// y = 1.  Multiplying by +-1 is exact, so s = (+-eta) + (+-x) * delta has the very same two roundings as
// +-(eta + x * delta): the sign is applied once per row instead of once per candidate.
template <> struct RowPair<CGG_BINOMIAL> {
    double e0, e1, x0, x1;
    bool z0, z1;     // y == 0: R's log(1 - p) branch (rform_log1p_rho)
    __device__ __forceinline__ RowPair(double2 y, double2 e, double2 x) {
        z0 = !(y.x > 0.5); z1 = !(y.y > 0.5);
        const double g0 = z0 ? 1.0 : -1.0, g1 = z1 ? 1.0 : -1.0;
        e0 = e.x * g0; e1 = e.y * g1; x0 = x.x * g0; x1 = x.y * g1;
    }
    __device__ __forceinline__ double term(double dk, double, const double2 *tab) const {
        double o0, o1;
        softplus2(eta_shift(e0, x0, dk), eta_shift(e1, x1, dk), tab, o0, o1, z0, z1);
        return -(o0 + o1);
    }
};

// mu = pmax(exp(eta'), eps); dpois_raw(y, mu) = y log(mu) - mu - lgamma(y + 1).  The lgamma term does not
// depend on beta and is added once per sum.  exp overflow gives -Inf as in R (!R_FINITE(lambda) -> R_D__0).
template <> struct RowPair<CGG_POISSON> {
    double y0, y1, e0, e1, x0, x1;
    __device__ __forceinline__ RowPair(double2 y, double2 e, double2 x) : y0(y.x), y1(y.y), e0(e.x), e1(e.y), x0(x.x), x1(x.y) {}
    __device__ __forceinline__ double term(double dk, double, const double2 *) const {
        double l0 = eta_shift(e0, x0, dk), l1 = eta_shift(e1, x1, dk);
        l0 = (l0 < kLogEps) ? kLogEps : l0; l1 = (l1 < kLogEps) ? kLogEps : l1;
        const double m0 = exp(l0), m1 = exp(l1);
        const double v0 = (m0 > 1.7976931348623157e308) ? -INFINITY : fma(y0, l0, -m0);
        const double v1 = (m1 > 1.7976931348623157e308) ? -INFINITY : fma(y1, l1, -m1);
        return v0 + v1;
    }
};

// Kernel-side family codes beyond the header's: binomial with the probit link is a family of its own here
constexpr int CGG_KF_NEGBIN = CGG_NEGATIVE_BINOMIAL;   // 3
constexpr int CGG_KF_PROBIT = 4;
constexpr double kProbitThresh = 8.125890664701906;    // -qnorm(.Machine$double.eps): stats' probit link clamps eta to +-this

// negative binomial, log link (R/glm_utils.R:55-57): dnbinom(y, size = 1, mu = pmax(exp(eta'), eps), log = TRUE)
//   = y log(mu) - (y + 1) log(1 + mu) = y l - (y + 1) softplus(l),  l = log(mu) = max(eta', log eps);  exp overflow -> -Inf
template <> struct RowPair<CGG_KF_NEGBIN> {
    double y0, y1, e0, e1, x0, x1;
    __device__ __forceinline__ RowPair(double2 y, double2 e, double2 x) : y0(y.x), y1(y.y), e0(e.x), e1(e.y), x0(x.x), x1(x.y) {}
    __device__ __forceinline__ double term(double dk, double, const double2 *tab) const {
        double l0 = eta_shift(e0, x0, dk), l1 = eta_shift(e1, x1, dk);
        l0 = (l0 < kLogEps) ? kLogEps : l0; l1 = (l1 < kLogEps) ? kLogEps : l1;
        double o0, o1;
        softplus2<false>(l0, l1, tab, o0, o1);
        const double v0 = (l0 > 709.782712893384) ? -INFINITY : fma(y0, l0, -(y0 + 1.0) * o0);
        const double v1 = (l1 > 709.782712893384) ? -INFINITY : fma(y1, l1, -(y1 + 1.0) * o1);
        return v0 + v1;
    }
};

// binomial, probit link (vignettes/pospkg.Rmd:88-108): dbinom(y, 1, p, log = TRUE) with p = pnorm(clamped eta'), q = 1 - p as
// R forms it; dbinom_raw's branches -bd0(1, p) - q (q < 0.1) and -bd0(1, q) - p (p < 0.1) equal log(p) = log1p(-q) and
// log(q) = log1p(-p) evaluated without cancellation.
__device__ __forceinline__ double probit_term(double y, double eta) {
    const double t = (eta < -kProbitThresh) ? -kProbitThresh : ((eta > kProbitThresh) ? kProbitThresh : eta);   // NaN stays NaN
    const double p = normcdf(t), q = 1.0 - p;
    if (y > 0.5) return (q < 0.1) ? log1p(-q) : log(p);
    return (p < 0.1) ? log1p(-p) : log(q);
}
template <> struct RowPair<CGG_KF_PROBIT> {
    double y0, y1, e0, e1, x0, x1;
    __device__ __forceinline__ RowPair(double2 y, double2 e, double2 x) : y0(y.x), y1(y.y), e0(e.x), e1(e.y), x0(x.x), x1(x.y) {}
    __device__ __forceinline__ double term(double dk, double, const double2 *) const {
        return probit_term(y0, eta_shift(e0, x0, dk)) + probit_term(y1, eta_shift(e1, x1, dk));
    }
};

// Single-row form (odd last row of a matrix, diagnostics): the pair with a neutral second row removed.
template <int FAMILY>
__device__ __forceinline__ double row_term(double y, double eta, double inv_sd, const double2 *tab) {
    if (FAMILY == CGG_GAUSSIAN) { const double z = (y - eta) * inv_sd; return -0.5 * z * z; }
    if (FAMILY == CGG_BINOMIAL) {
        double o0, o1;
        const double s = (y > 0.5) ? -eta : eta;
        softplus2(s, s, tab, o0, o1, !(y > 0.5), false);
        return -o0;
    }
    if (FAMILY == CGG_KF_PROBIT) return probit_term(y, eta);
    if (FAMILY == CGG_KF_NEGBIN) {
        const double l = (eta < kLogEps) ? kLogEps : eta;
        double o0, o1;
        softplus2<false>(l, l, tab, o0, o1);
        return (l > 709.782712893384) ? -INFINITY : fma(y, l, -(y + 1.0) * o0);
    }
    double le = (eta < kLogEps) ? kLogEps : eta;
    const double mu = exp(le);
    return (mu > 1.7976931348623157e308) ? -INFINITY : fma(y, le, -mu);
}

// Cooperative copy of the math tables into shared memory (call from every thread, then sync): the log1p split table
// (L1P_N + 1 double2 entries) followed by the 64 doubles 2^(j/64) of the light jet pass's exp (cgg_jet.cuh).
constexpr int MATH_TAB_N = L1P_N + 1 + EX64_N / 2;
__device__ __forceinline__ void load_l1p_table(double2 *dst) {
    for (int i = threadIdx.x; i <= L1P_N; i += blockDim.x) dst[i] = make_double2(L1P_TAB[2 * i], L1P_TAB[2 * i + 1]);
    for (int i = threadIdx.x; i < EX64_N / 2; i += blockDim.x) dst[L1P_N + 1 + i] = make_double2(EX64_TAB[2 * i], EX64_TAB[2 * i + 1]);
}
__device__ __forceinline__ const double *ex64_table(const double2 *tab) { return reinterpret_cast<const double *>(tab + L1P_N + 1); }

struct PriorParams {
    int kind;
    double mu, sigma, df;
    double c0;       // additive constant of one coordinate's log-density
    double inv_sigma;
};

// distributional::density(beta_prior, x, log = TRUE) for one coordinate (R/glm_utils.R:109)
__device__ __forceinline__ double prior_logdens1(const PriorParams &pp, double x) {
    double z = (x - pp.mu) * pp.inv_sigma;
    if (pp.kind == CGG_PRIOR_NORMAL) return pp.c0 - 0.5 * z * z;        // -(ln sqrt(2pi) + z^2/2 + log sigma)
    if (pp.kind == CGG_PRIOR_LAPLACE) return pp.c0 - fabs(z);           // -log(2 sigma) - |x - mu| / sigma
    if (pp.kind == CGG_PRIOR_STUDENT_T) return pp.c0 - 0.5 * (pp.df + 1.0) * log1p(z * z / pp.df);   // dt(z, df, log=TRUE) - log(sigma)
    if (x != x) return x;
    if (pp.kind == CGG_PRIOR_GAMMA) {                                   // dgamma(x, shape = mu, rate = sigma, log = TRUE); c0 = shape log(rate) - lgamma(shape)
        if (x < 0.0) return -INFINITY;
        if (x == 0.0) return (pp.mu < 1.0) ? INFINITY : ((pp.mu > 1.0) ? -INFINITY : pp.c0);
        return pp.c0 + (pp.mu - 1.0) * log(x) - pp.sigma * x;
    }
    return (x < 0.0) ? -INFINITY : pp.c0 - pp.sigma * x;                // dexp(x, rate = sigma, log = TRUE); c0 = log(rate)
}
// A list of priors: every component at every coordinate, summed (log_prior_density.list, R/glm_utils.R:113-115, quirk Q6)
struct PriorSet {
    PriorParams comp[CGG_MAX_PRIORS];
    int n, pad;
};
__device__ __forceinline__ double prior_logdens(const PriorSet &ps, double x) {
    double v = prior_logdens1(ps.comp[0], x);
    for (int k = 1; k < ps.n; ++k) v += prior_logdens1(ps.comp[k], x);
    return v;
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
