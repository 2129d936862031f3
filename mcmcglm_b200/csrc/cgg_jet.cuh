// cgg_jet.cuh -- the "jet" pass: one pass over (y, eta, X_j) that delivers the exact log-likelihood at the
// chain's current point AND the derivatives of the log-likelihood along coordinate j, so that every candidate
// the slice sampler will ask about can be enclosed by the decider without touching the rows again.
//
// For a candidate b the reference evaluates (R/glm_utils.R:187-218)
//     f(b) = sum_i l(eta_i + x_ij * delta; y_i) + log-prior,   delta = b - beta_j,
// with l the per-row log-density of the family.  Along the coordinate this is a smooth scalar function of
// delta, so with xs_i = x_ij * cs (cs a power of two chosen per column such that max|xs_i| <= 1: exact scaling)
// and h = delta / cs
//     sum_i l(eta_i + xs_i h) = sum_{k=0..D} M_k h^k / k!  +  R_D(h),     M_k = sum_i xs_i^k l^(k)(eta_i; y_i),
//     |R_D(h)| <= |h|^(D+1)/(D+1)! * sum_i |xs_i|^(D+1) sup|l^(D+1)|.
// The pass accumulates M_0..M_D (M_0 with the row routine the exact passes use -- poisson: the same expression with a
// branch-free exp that agrees with libdevice's to ~2 ulp of mu, inside the bound's rounding envelope -- so it is the
// f(x0) the reference's first evaluation returns) plus the few sums the error bound needs.  jet_eval() returns
// the surrogate value and a bound B on |surrogate - what an exact fp64 pass would return|; the decider
// accepts / rejects a candidate from the surrogate only when the comparison with the slice level holds with
// margin B, and asks for an exact pass otherwise.  Results are therefore those of the all-exact engine.
//
// Slot layout of the NV values a jet pass delivers:
//   gaussian : 0..2 M_0..M_2 (exact quadratic, no remainder) | 3 sum|z||eta| | 4 sum|z| | 5 sum|eta|
//   binomial : 0 M_0 (full passes only; light passes leave it 0) | 1..7 the positive-form sums m_1..m_7 (see JetRow<CGG_BINOMIAL>) |
//              8 an upper bound of sum_i e^|eta_i| (rform_noise_sum: bounds the rounding of p that R's log(1 - p) carries) |
//              9 rows with |eta| >= 21.9 (the stats logit clamp at |eta| = 30 is then within reach of the enclosure's radius)
//   poisson  : 0..6 M_0..M_6 | 7 sum|xs|^7 mu | 8 sum(|y| + mu)(|eta| + 1) | 9 rows too close to the pmax(., eps) clamp
#pragma once
#include "cgg_math.cuh"

namespace cgg {

#ifndef CGG_JET_D
#define CGG_JET_D 5                        // order of the binomial expansion (5 or 7)
#endif
constexpr int JET_NV = CGG_KMAX + 2;       // same number of accumulators as a candidate pass
constexpr int CS_STRIDE = 12;              // per-column statistics: {cs, 1/cs, S_1..S_8, max|x|, C1 = sum xs (y - 1/2)}
constexpr double JET_AMAX = 8.0;           // the enclosure is only used for |h| <= JET_AMAX (binomial, poisson)
constexpr double JET_EPS = 1.1102230246251565e-16;   // 2^-53
constexpr double JET_CROUND = 256.0;       // least rounding allowance of one accumulated moment, in units of eps * sum|terms|: the engine
                                           // passes jet_eval the larger of this and what the summation depth of the run needs (Dev::jet_ce:
                                           // rows per lane + the warp / CTA / grid folds; n > ~1.4e7 rows per GPU exceed 256)
constexpr int JET_DL = 3;                  // order of the binomial LIGHT pass
constexpr double JET_LIGHT_EPS = 2e-11;    // absolute error of sigmoid-derived quantities (ua, v, v ua) in a light pass: the reciprocal
                                           // seed is taken from the high word of 1 + T (>= 2^-19.5 relative), one Newton step squares it;
                                           // measured by cgg_debug_light_error (tests/test_gpu_jet.py asserts <= half of this)

// sup_t |softplus^(k)(t)|, k = 1..8, rounded up (tools/gen_math_tables.py prints them; polynomial in sigmoid)
__device__ __constant__ double JET_G[9] = {0.0, 1.0, 0.25, 0.0962250449, 0.125, 0.127683922, 0.25, 0.408327759, 1.0625};
// 1/k!
__device__ __constant__ double JET_IFACT[9] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0};

// ---- per-row accumulation ---------------------------------------------------------------------------
// m[] are per-lane running sums; (y, e, xs) one row with e the committed linear predictor and xs the scaled x.
template <int FAMILY> struct JetRow;

template <> struct JetRow<CGG_GAUSSIAN> {
    static constexpr unsigned RISK_KEY = 0xffffffffu;
    template <bool FULL>
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *,
                                                double (&m)[JET_NV], unsigned &) {
        // M_0 with the exact pass's own expression (RowPair<GAUSSIAN>::term at delta = 0)
        const double z0 = (y.x - e.x) * inv_sd, z1 = (y.y - e.y) * inv_sd;
        m[0] += -0.5 * fma(z0, z0, z1 * z1);
        const double g0 = xs.x * inv_sd, g1 = xs.y * inv_sd;
        m[1] += fma(g0, z0, g1 * z1);                    // d/dh: z * xs / sd
        m[2] -= fma(g0, g0, g1 * g1);                    // d2/dh2: -(xs / sd)^2
        const double az0 = fabs(z0), az1 = fabs(z1), ae0 = fabs(e.x), ae1 = fabs(e.y);
        m[3] += fma(az0, ae0, az1 * ae1);
        m[4] += az0 + az1;
        m[5] += ae0 + ae1;
    }
    template <bool FULL>
    static __device__ __forceinline__ void add1(double y, double e, double xs, double inv_sd, const double2 *, double (&m)[JET_NV], unsigned &) {
        const double z = (y - e) * inv_sd, g = xs * inv_sd;
        m[0] += -0.5 * z * z; m[1] += g * z; m[2] -= g * g;
        m[3] += fabs(z) * fabs(e); m[4] += fabs(z); m[5] += fabs(e);
    }
};

// binomial-logit: l(t) = y t - softplus(t).  Everything is evaluated at a = |eta| >= 0 with ONE exp(-a):
//   M_0 term  -softplus(s), s = +-eta by y (the exact passes' own expression: relu(s) + log1p(exp(-a)));
//   l'(eta)   = y - sigmoid(eta) = (y - 1/2) - sign(eta) ua / 2,         ua = 2 sigmoid(a) - 1 = tanh(a/2) >= 0;
//   l^(k)(eta) = -sign(eta)^k sp^(k)(a) for k >= 2 (sp2 is even, sp3 odd, ...), polynomials in v = s(1-s) and ua:
//       sp2 = v, sp3 = -v ua, sp4 = v(1-6v), sp5 = -v ua (1-12v), sp6 = v(1-30v+120v^2), sp7 = -v ua (1-60v+360v^2).
// With xh = sign(eta) * xs (a sign-bit XOR) the pass accumulates the all-positive-form sums
//   m1 = sum xh ua, m2 = sum xh^2 v, m3 = sum xh^3 v ua, m4 = sum xh^4 v(1-6v), ...      and jet_eval() forms
//   M_1 = C1 - m1/2 (C1 = sum xs (y - 1/2), a per-column constant), M_2 = -m2, M_3 = +m3, M_4 = -m4, M_5 = +m5, ...
// No fp64 select or compare besides the clamp: signs are handled on the high words with integer logic, which
// runs on the ALU pipe while the fp64 pipe is the bottleneck.  1/(1+T) = tab.x / (1+w) with the log1p split
// T = c + w (1+c), |w| <= 1/64: (1-w)(1+w^2) and one Newton step (absolute error < 4e-15, inside JET_CROUND).
template <> struct JetRow<CGG_BINOMIAL> {
    static constexpr unsigned RISK_KEY = 0x4035e666u;     // |eta| >= 21.9: the clamp at |eta| = 30 is within JET_AMAX
    // FULL: also the exact M_0 term (softplus through the log1p table, stats' clamp).  Light passes skip it: the slice
    // decisions only involve differences f(v) - f(x0), in which M_0 cancels; 1/(1+T) then comes from the hardware
    // reciprocal seed (rcp.approx.ftz.f64, 2^-23) and two Newton steps instead of the table split.
    // LIGHT passes (the product path in the stationary regime): order JET_DL = 3 and every per-row quantity only to
    // JET_LIGHT_EPS absolute -- the approximation errors enter the enclosure's bound (jet_eval), which stays ~1e-7 against
    // slice margins of O(1), and a test the looser enclosure cannot decide is repeated as a FULL pass (order CGG_JET_D,
    // 1-ulp routines, exact M_0).  exp(-a) = 2^(K/64) e^r: K = rint(-a 64/ln2), one-constant reduction, 64-entry table in
    // shared memory, degree-4 polynomial on |r| <= ln2/128 (1.8e-14 relative, tools/gen_math_tables.py); 1/(1+T) from the
    // hardware seed and ONE Newton step.  20 fp64 instructions per row (FULL: ~60; round 1's light pass: 40).
    static __device__ __forceinline__ void add_light(double e, double xs, const double2 *tab, double (&m)[JET_NV], unsigned &risk) {
        const double SHIFT = 6755399441055744.0;
        const int he = __double2hiint(e);
        risk = max(risk, (unsigned)he & 0x7fffffffu);
        const double a = fabs(e);
        const double kd = fma(-a, 9.2332482616893656e+01 /* 64 / ln2 */, SHIFT);
        const double r = fma(kd - SHIFT, -1.0830424696249145e-02 /* ln2 / 64 */, -a);
        double p = fma(EX64_C[4], r, EX64_C[3]);
        p = fma(p, r, EX64_C[2]);
        p = fma(p, r, 1.0);
        p = fma(p, r, 1.0);
        const int K = __double2loint(kd);      // <= 0; meaningless for |eta| >~ 1e7, far beyond the risk key: T and the sums may then
                                               // be anything, NaN included -- the pass is flagged (m[9]) and its sums are not used
        const double T = scale2(ex64_table(tab)[K & (EX64_N - 1)] * p, K >> 6);   // exp(-a)
        const double dd = 1.0 + T;                                           // in (1, 2]
        double rr;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(dd));
        rr = fma(rr, fma(-dd, rr, 1.0), rr);                                 // sigmoid(a)
        const double v = fma(-rr, rr, rr);
        const double ua = fma(2.0, rr, -1.0);
        const double xh = __hiloint2double(__double2hiint(xs) ^ (he & 0x80000000), __double2loint(xs));
        const double x2 = xh * xh;
        m[1] = fma(xh, ua, m[1]);
        m[2] = fma(x2, v, m[2]);
        m[3] = fma(x2 * xh, v * ua, m[3]);
    }
    template <bool FULL>
    static __device__ __forceinline__ void add1(double y, double e, double xs, double, const double2 *tab, double (&m)[JET_NV], unsigned &risk) {
        if (!FULL) { add_light(e, xs, tab, m, risk); return; }
        const double SHIFT = 6755399441055744.0;
        const int he = __double2hiint(e);
        risk = max(risk, (unsigned)he & 0x7fffffffu);                       // max |eta| (high word): checked against 21.9 at the end
        double a = fabs(e);
        if (FULL) a = (a > 30.0) ? kLogitClampEta : a;
        const double kd = fma(-a, 1.4426950408889634, SHIFT);
        const double kf = kd - SHIFT;
        double r = fma(kf, -6.93147180369123816490e-01, -a);
        r = fma(kf, -1.90821492927058770002e-10, r);
        double pe, po;
        poly_exp_eo(r, pe, po);
        const double p = fma(po, r, pe);
        const double T = scale2(p, __double2loint(kd));                      // exp(-a)
        double rr;
        if (FULL) {
            // y = 0 and 8 < eta <= 30: R's log(1 - p) carries the rounding of p (cgg_math.cuh, rform_log1p_rho)
            if (!(y > 0.5) && e > kRFormLo && e <= 30.0) m[0] += rform_log1p_rho(T, scale2(fma(-po, r, pe), -__double2loint(kd)));
            const int hs = he ^ ((__double2hiint(y) << 2) & 0x80000000);    // sign of s = (y == 1) ? -eta : eta
            const double md = fma(T, (double)L1P_N, SHIFT);
            int mi = __double2loint(md);
            mi = min(max(mi, 0), L1P_N);
            const double2 tb = tab[mi];
            const double w = fma(md - SHIFT, -1.0 / L1P_N, T) * tb.x;
            const double w2 = w * w;
            const double q = poly_l1p_q(w, w2);
            const double l1p = tb.y + fma(w2, q, w);
            const int keep = ~(hs >> 31);                                    // all ones iff s >= 0
            const double relu = __hiloint2double(__double2hiint(a) & keep, __double2loint(a) & keep);
            m[0] -= relu + l1p;
            const double omw = 1.0 - w;
            rr = fma(w2, omw, omw);
            rr = fma(rr, fma(-(1.0 + w), rr, 1.0), rr);
            rr *= tb.x;                                                      // sigmoid(a)
        } else {
            const double dd = 1.0 + T;                                       // in (1, 2]
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(dd));
            rr = fma(rr, fma(-dd, rr, 1.0), rr);
            rr = fma(rr, fma(-dd, rr, 1.0), rr);
        }
        const double v = fma(-rr, rr, rr);                                   // s(1-s), absolute error ~eps
        const double ua = fma(2.0, rr, -1.0);                                // 2 sigmoid(a) - 1
        const double xh = __hiloint2double(__double2hiint(xs) ^ (he & 0x80000000), __double2loint(xs));
        const double vu = v * ua;
        const double x2 = xh * xh, x3 = x2 * xh, x4 = x2 * x2, x5 = x4 * xh;
        m[1] = fma(xh, ua, m[1]);
        m[2] = fma(x2, v, m[2]);
        m[3] = fma(x3, vu, m[3]);
        m[4] = fma(x4, v * fma(-6.0, v, 1.0), m[4]);
        m[5] = fma(x5, vu * fma(-12.0, v, 1.0), m[5]);
#if CGG_JET_D >= 7
        const double x6 = x3 * x3, x7 = x6 * xh;
        m[6] = fma(x6, v * fma(fma(120.0, v, -30.0), v, 1.0), m[6]);
        m[7] = fma(x7, vu * fma(fma(360.0, v, -60.0), v, 1.0), m[7]);
#endif
    }
    template <bool FULL>
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *tab, double (&m)[JET_NV], unsigned &risk) {
        add1<FULL>(y.x, e.x, xs.x, inv_sd, tab, m, risk);
        add1<FULL>(y.y, e.y, xs.y, inv_sd, tab, m, risk);
    }
};

// poisson-log: l(t) = y t - exp(t) (- lgamma(y+1), per-dataset constant); l' = y - mu, l^(k) = -mu for k >= 2.
template <> struct JetRow<CGG_POISSON> {
    static constexpr unsigned RISK_KEY = 0xc03be666u;     // eta <= -27.9: the pmax(., eps) clamp at -36.04 is within JET_AMAX
    template <bool FULL>
    static __device__ __forceinline__ void add1(double y, double e, double xs, double, const double2 *, double (&m)[JET_NV], unsigned &risk) {
        const double l = (e < kLogEps) ? kLogEps : e;
        // exp(l), l >= -36.04: Cody-Waite reduction + the degree-11 polynomial of the fused softplus (<= 1 ulp, branch-free,
        // ~17 fp64 instructions against ~30 for the general-purpose exp); 2^k by exponent arithmetic, overflow -> +Inf
        const double SHIFT = 6755399441055744.0;
        const double kd = fma(l, 1.4426950408889634, SHIFT);
        const double kf = kd - SHIFT;
        double r = fma(kf, -6.93147180369123816490e-01, l);
        r = fma(kf, -1.90821492927058770002e-10, r);
        const double pe = poly_exp(r);
        double mu = __hiloint2double(__double2hiint(pe) + (__double2loint(kd) << 20), __double2loint(pe));
        mu = (l > 709.0) ? INFINITY : mu;
        m[0] += (mu > 1.7976931348623157e308) ? -INFINITY : fma(y, l, -mu);      // row_term<POISSON> (same value within 2 ulp of mu)
        const double nm = -mu;
        const double x2 = xs * xs, x3 = x2 * xs, x4 = x2 * x2, x5 = x4 * xs, x6 = x3 * x3, x7 = x6 * xs;
        m[1] = fma(xs, y - mu, m[1]);
        m[2] = fma(x2, nm, m[2]);
        m[3] = fma(x3, nm, m[3]);
        m[4] = fma(x4, nm, m[4]);
        m[5] = fma(x5, nm, m[5]);
        m[6] = fma(x6, nm, m[6]);
        m[7] = fma(fabs(x7), mu, m[7]);
        const double ae = fabs(e);
        m[8] = fma(fabs(y) + mu, ae + 1.0, m[8]);
        risk = max(risk, (unsigned)__double2hiint(e));                     // most negative eta (high word): checked against -27.9
    }
    template <bool FULL>
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *tab, double (&m)[JET_NV], unsigned &risk) {
        add1<FULL>(y.x, e.x, xs.x, inv_sd, tab, m, risk);
        add1<FULL>(y.y, e.y, xs.y, inv_sd, tab, m, risk);
    }
};

// negative binomial and binomial-probit: exact passes only (the engine never asks for a jet pass); stubs so that the pass
// templates instantiate
template <int FAMILY> struct JetRowNone {
    static constexpr unsigned RISK_KEY = 0u;       // "every row is a risk": an enclosure would never apply
    template <bool FULL>
    static __device__ __forceinline__ void add1(double, double, double, double, const double2 *, double (&)[JET_NV], unsigned &) {}
    template <bool FULL>
    static __device__ __forceinline__ void add2(double2, double2, double2, double, const double2 *, double (&)[JET_NV], unsigned &) {}
};
template <> struct JetRow<CGG_KF_NEGBIN> : JetRowNone<CGG_KF_NEGBIN> {};
template <> struct JetRow<CGG_KF_PROBIT> : JetRowNone<CGG_KF_PROBIT> {};

// A binomial LIGHT pass delivers five values, compactly: m1, m2, m3, the noise sum (canonical slot 8), the risk-row count
// (canonical slot 9).  pack: canonical -> compact (worker, before the reduction); unpack: compact -> canonical (decider).
constexpr int JET_NVL = 5;
__device__ __forceinline__ int jet_nvals(int family, bool light) { return (family == CGG_BINOMIAL && light) ? JET_NVL : JET_NV; }
__device__ __forceinline__ void jet_light_pack(double (&m)[JET_NV]) { m[0] = m[1]; m[1] = m[2]; m[2] = m[3]; m[3] = m[8]; m[4] = m[9]; }
__device__ __forceinline__ void jet_light_unpack(double (&m)[JET_NV]) {
    m[9] = m[4]; m[8] = m[3]; m[3] = m[2]; m[2] = m[1]; m[1] = m[0];
    m[0] = 0.0; m[4] = 0.0; m[5] = 0.0; m[6] = 0.0; m[7] = 0.0;
}

// Upper bound of sum_i e^|eta_i| over the `rows` rows a lane scored, from the running maximum of the high words of
// |eta_i| (JetRow<CGG_BINOMIAL>'s risk key): rows * e^amax, amax rounded up, ex2.approx (2^-22) and the fp32 roundings
// covered by the factor 1.001.  Overflow gives +Inf and a NaN key NaN: the enclosure then decides nothing.
__device__ __forceinline__ double rform_noise_sum(unsigned risk_key, unsigned rows) {
    const float amax = __double2float_ru(__hiloint2double((int)risk_key, -1));
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(__fmul_ru(amax, 1.44269514f)));
    return (double)rows * (double)(ex * 1.001f);
}

// ---- the enclosure --------------------------------------------------------------------------------
// m[]: the pass's sums; cst: the column's statistics; delta = cand - x0; fmag: the magnitude of the log-likelihood
// (|M_0|, or the carried |f(x0)| after a light pass).  Returns the surrogate log-likelihood DIFFERENCE
// sum_{k>=1} M_k h^k / k! and in B a bound on its distance from the difference (f(cand) - f(x0)) of two exact fp64
// evaluations of this engine: Taylor remainder + rounding of the accumulated moments + the rounding envelopes of the
// two exact evaluations.  B is +Inf (or NaN) when the enclosure does not apply: the caller must treat any comparison
// that is not strictly decided as undecided.
__device__ __forceinline__ double jet_eval(int family, const double (&m)[JET_NV], const double *cst, double n, double inv_sd,
                                           double delta, double fmag, double &B, bool light = false, double ce_in = 0.0) {
    const double h = delta * cst[1];
    const double a = fabs(h);
    const double ce = fmax(ce_in, JET_CROUND * JET_EPS);
    if (family == CGG_GAUSSIAN) {
        const double dl = h * fma(0.5 * h, m[2], m[1]);
        // moments: sum|xs z|/sd <= sqrt(S_2 * 2|M_0|)/sd, |M_2|
        const double bmom = ce * (a * sqrt(cst[3] * 2.0 * fabs(m[0])) * inv_sd + 0.5 * a * a * fabs(m[2]));
        // exact pass: t = fl(eta + fl(x delta)) perturbs -z^2/2 by <= |z| |dt| / sd, the rest is relative to z^2
        const double bex = JET_EPS * inv_sd * (m[3] + 2.0 * a * m[4] + a * inv_sd * (m[5] + 2.0 * a * n)) + ce * (2.0 * fmag + fabs(dl));
        B = 1.01 * (bmom + bex);
        return dl;
    }
    if (family == CGG_BINOMIAL) {
        constexpr int DF = CGG_JET_D;
        static_assert((DF & 1) && (JET_DL & 1) && JET_DL <= DF, "odd orders: the top moment enters with a + sign");
        const int D = light ? JET_DL : DF;
        // signed moments from the positive-form sums: M_1 = C1 - m1/2, M_k = (-1)^(k+1) m_k
        double f = 0.0;
#pragma unroll
        for (int k = DF; k >= 2; --k) if (k <= D) f = fma(f, h, ((k & 1) ? m[k] : -m[k]) * JET_IFACT[k]);
        f = fma(f, h, fma(-0.5, m[1], cst[11]));
        const double dl = f * h;
        // remainder: G_(D+1) S_(D+1) a^(D+1) / (D+1)!;  moment k: rounding <= ce G_k S_k (terms are bounded by |xs|^k G_k);
        // light passes: + the approximation error of the row quantities, JET_LIGHT_EPS |xs|^k per term
        double pw = a, bmom = 0.0, bapx = 0.0;
#pragma unroll
        for (int k = 1; k <= DF; ++k) if (k <= D) { bmom = fma(JET_G[k] * JET_IFACT[k] * cst[1 + k], pw, bmom); bapx = fma(JET_IFACT[k] * cst[1 + k], pw, bapx); pw *= a; }
        double bt = JET_G[DF + 1] * JET_IFACT[DF + 1] * cst[2 + DF] * pw;
        if (light) bt = JET_G[JET_DL + 1] * JET_IFACT[JET_DL + 1] * cst[2 + JET_DL] * pw + JET_LIGHT_EPS * bapx;
        // exact passes: |l'| <= 1, so the rounding of t costs <= eps (|eta| + 2 |x delta|) with every |eta| < 21.9 (else
        // m[9] != 0); softplus and sums relative to |f|, for both evaluations
        const double bex = JET_EPS * (21.9 * n + 2.0 * a * cst[2]) + ce * (2.0 * fmag + fabs(dl) + bt);
        // R's log(1 - p) form (rform_log1p_rho): an exact evaluation deviates from the smooth function by at most
        // 2^-54 (1 + e^t) per row, t <= |eta| + a; both evaluations of the difference, and slack for log1p(rho) vs rho
        float ea;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(__fmul_ru(__double2float_ru(a), 1.44269514f)));
        const double brf = 2.0 * JET_EPS * (n + (double)(ea * 1.001f) * m[8]);
        B = 1.01 * (bt + ce * bmom + bex + brf);
        if (!(a <= JET_AMAX) || m[9] != 0.0) B = INFINITY;
        return dl;
    }
    double f = m[6] * JET_IFACT[6];
#pragma unroll
    for (int k = 5; k >= 1; --k) f = fma(f, h, m[k] * JET_IFACT[k]);
    const double dl = f * h;
    const double ea = exp(a);
    const double a2 = a * a, a4 = a2 * a2;
    const double bt = ea * (a4 * a2 * a) * JET_IFACT[7] * m[7] * (1.0 + 1e-6);
    // every term of every moment, of the exact passes and of their t-rounding is bounded by e^a (1 + 2a) (|y| + mu)(|eta| + 1)
    const double brnd = (ce + 100.0 * JET_EPS) * ea * (1.0 + 2.0 * a) * m[8];
    B = 1.01 * (bt + brnd);
    if (!(a <= JET_AMAX) || m[9] != 0.0) B = INFINITY;
    return dl;
}

}  // namespace cgg
