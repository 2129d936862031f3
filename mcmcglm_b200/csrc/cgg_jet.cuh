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
// The pass accumulates M_0..M_D (M_0 with the very row routine the exact passes use, so it IS the exact
// f(x0) the reference's first evaluation returns) plus the few sums the error bound needs.  jet_eval() returns
// the surrogate value and a bound B on |surrogate - what an exact fp64 pass would return|; the decider
// accepts / rejects a candidate from the surrogate only when the comparison with the slice level holds with
// margin B, and asks for an exact pass otherwise.  Results are therefore those of the all-exact engine.
//
// Slot layout of the NV values a jet pass delivers:
//   gaussian : 0..2 M_0..M_2 (exact quadratic, no remainder) | 3 sum|z||eta| | 4 sum|z| | 5 sum|eta|
//   binomial : 0..7 M_0..M_7 | 8 sum|eta| | 9 rows too close to the stats logit clamp (|eta| = 30)
//   poisson  : 0..6 M_0..M_6 | 7 sum|xs|^7 mu | 8 sum(|y| + mu)(|eta| + 1) | 9 rows too close to the pmax(., eps) clamp
#pragma once
#include "cgg_math.cuh"

namespace cgg {

constexpr int JET_NV = CGG_KMAX + 2;       // same number of accumulators as a candidate pass
constexpr int CS_STRIDE = 12;              // per-column statistics: {cs, 1/cs, S_1..S_8, max|x|, pad}
constexpr double JET_AMAX = 8.0;           // the enclosure is only used for |h| <= JET_AMAX (binomial, poisson)
constexpr double JET_EPS = 1.1102230246251565e-16;   // 2^-53
constexpr double JET_CROUND = 64.0;        // rounding allowance of one accumulated moment, in units of eps * sum|terms|

// sup_t |softplus^(k)(t)|, k = 1..8, rounded up (tools/gen_math_tables.py prints them; polynomial in sigmoid)
__device__ __constant__ double JET_G[9] = {0.0, 1.0, 0.25, 0.0962250449, 0.125, 0.127683922, 0.25, 0.408327759, 1.0625};
// 1/k!
__device__ __constant__ double JET_IFACT[9] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0};

// ---- per-row accumulation ---------------------------------------------------------------------------
// m[] are per-lane running sums; (y, e, xs) one row with e the committed linear predictor and xs the scaled x.
template <int FAMILY> struct JetRow;

template <> struct JetRow<CGG_GAUSSIAN> {
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *,
                                                double (&m)[JET_NV]) {
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
    static __device__ __forceinline__ void add1(double y, double e, double xs, double inv_sd, const double2 *, double (&m)[JET_NV]) {
        const double z = (y - e) * inv_sd, g = xs * inv_sd;
        m[0] += -0.5 * z * z; m[1] += g * z; m[2] -= g * g;
        m[3] += fabs(z) * fabs(e); m[4] += fabs(z); m[5] += fabs(e);
    }
};

// binomial-logit: l(t) = y t - softplus(t); l' = y - s, l^(k) = -softplus^(k) (k >= 2), all polynomials in
// s = sigmoid(t): with v = s(1-s), u = 1-2s:  sp2 = v, sp3 = v u, sp4 = v(1-6v), sp5 = v u (1-12v),
// sp6 = v(1-30v+120v^2), sp7 = v u (1-60v+360v^2).  One exp(-|t|) feeds both the softplus (M_0, same code as
// softplus2) and the sigmoid: 1/(1+T) = tab.x / (1+w) with the log1p split T = c + w(1+c), |w| <= 1/64.
struct BinomJetPieces { double sp, s, v, u; };
__device__ __forceinline__ BinomJetPieces binom_pieces(double sarg /* +-eta as softplus2 gets it */, double eta, const double2 *tab) {
    const double SHIFT = 6755399441055744.0;
    double a = fabs(sarg);
    a = (a > 30.0) ? kLogitClampEta : a;
    const double kd = fma(-a, 1.4426950408889634, SHIFT);
    const double kf = kd - SHIFT;
    double r = fma(kf, -6.93147180369123816490e-01, -a);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double p = poly_exp(r);
    const double T = __hiloint2double(__double2hiint(p) + (__double2loint(kd) << 20), __double2loint(p));   // exp(-a)
    const double md = fma(T, (double)L1P_N, SHIFT);
    int mi = __double2loint(md);
    mi = min(max(mi, 0), L1P_N);
    const double2 tb = tab[mi];
    const double w = fma(md - SHIFT, -1.0 / L1P_N, T) * tb.x;
    const double w2 = w * w;
    const double q = poly_l1p_q(w, w2);
    const double l1p = tb.y + fma(w2, q, w);
    BinomJetPieces o;
    o.sp = ((sarg > 0.0) ? a : 0.0) + l1p;
    // 1/(1+w): (1-w)(1+w^2) then two Newton steps (error w^4 -> w^8 -> w^16)
    const double omw = 1.0 - w, opw = 1.0 + w;
    double rr = fma(w2, omw, omw);
    rr = fma(rr, fma(-opw, rr, 1.0), rr);
    rr = fma(rr, fma(-opw, rr, 1.0), rr);
    rr *= tb.x;                                   // 1/(1+T) = sigmoid(a)
    const double sneg = T * rr;                   // sigmoid(-a)
    o.s = (eta > 0.0) ? rr : sneg;                // sigmoid(eta)
    o.v = sneg * rr;                              // s (1 - s)
    const double ua = (1.0 - T) * rr;             // |1 - 2 s|
    o.u = (eta > 0.0) ? -ua : ua;
    return o;
}

template <> struct JetRow<CGG_BINOMIAL> {
    static __device__ __forceinline__ void add1(double y, double e, double xs, double, const double2 *tab, double (&m)[JET_NV]) {
        const double sarg = (y > 0.5) ? -e : e;
        const BinomJetPieces b = binom_pieces(sarg, e, tab);
        m[0] -= b.sp;
        const double q2 = -b.v, qu = q2 * b.u;
        const double a4 = fma(-6.0, b.v, 1.0), a5 = fma(-12.0, b.v, 1.0);
        const double a6 = fma(fma(120.0, b.v, -30.0), b.v, 1.0), a7 = fma(fma(360.0, b.v, -60.0), b.v, 1.0);
        const double x2 = xs * xs, x3 = x2 * xs, x4 = x2 * x2, x5 = x4 * xs, x6 = x3 * x3, x7 = x6 * xs;
        m[1] = fma(xs, y - b.s, m[1]);
        m[2] = fma(x2, q2, m[2]);
        m[3] = fma(x3, qu, m[3]);
        m[4] = fma(x4, q2 * a4, m[4]);
        m[5] = fma(x5, qu * a5, m[5]);
        m[6] = fma(x6, q2 * a6, m[6]);
        m[7] = fma(x7, qu * a7, m[7]);
        const double ae = fabs(e);
        m[8] += ae;
        m[9] += (fma(fabs(xs), JET_AMAX, ae) >= 29.9) ? 1.0 : 0.0;
    }
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *tab, double (&m)[JET_NV]) {
        add1(y.x, e.x, xs.x, inv_sd, tab, m);
        add1(y.y, e.y, xs.y, inv_sd, tab, m);
    }
};

// poisson-log: l(t) = y t - exp(t) (- lgamma(y+1), per-dataset constant); l' = y - mu, l^(k) = -mu for k >= 2.
template <> struct JetRow<CGG_POISSON> {
    static __device__ __forceinline__ void add1(double y, double e, double xs, double, const double2 *, double (&m)[JET_NV]) {
        const double l = (e < kLogEps) ? kLogEps : e;
        const double mu = exp(l);
        m[0] += (mu > 1.7976931348623157e308) ? -INFINITY : fma(y, l, -mu);      // row_term<POISSON>
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
        m[9] += (fma(-fabs(xs), JET_AMAX, e) <= kLogEps + 0.1) ? 1.0 : 0.0;
    }
    static __device__ __forceinline__ void add2(double2 y, double2 e, double2 xs, double inv_sd, const double2 *tab, double (&m)[JET_NV]) {
        add1(y.x, e.x, xs.x, inv_sd, tab, m);
        add1(y.y, e.y, xs.y, inv_sd, tab, m);
    }
};

// ---- the enclosure --------------------------------------------------------------------------------
// m[]: the pass's sums; cst: the column's statistics; delta = cand - x0.  Returns the surrogate
// log-likelihood (without ll_const and without the prior) and in B a bound on its distance from the value an
// exact fp64 pass of this engine would deliver (Taylor remainder + rounding of the accumulated moments +
// rounding envelope of the exact evaluation itself).  B is +Inf (or NaN) when the enclosure does not apply:
// the caller must treat any comparison that is not strictly decided as undecided.
__device__ __forceinline__ double jet_eval(int family, const double (&m)[JET_NV], const double *cst, double n, double inv_sd,
                                           double delta, double &B) {
    const double h = delta * cst[1];
    const double a = fabs(h);
    const double ce = JET_CROUND * JET_EPS;
    if (family == CGG_GAUSSIAN) {
        const double f = fma(h, fma(0.5 * h, m[2], m[1]), m[0]);
        // moments: |M_0| (same-sign terms), sum|xs z|/sd <= sqrt(S_2 * 2|M_0|)/sd, |M_2|
        const double bmom = ce * (fabs(m[0]) + a * sqrt(cst[3] * 2.0 * fabs(m[0])) * inv_sd + 0.5 * a * a * fabs(m[2]));
        // exact pass: t = fl(eta + fl(x delta)) perturbs -z^2/2 by <= |z| |dt| / sd, the rest is relative to z^2
        const double bex = JET_EPS * inv_sd * (m[3] + 2.0 * a * m[4] + a * inv_sd * (m[5] + 2.0 * a * n)) + ce * fabs(f);
        B = 1.01 * (bmom + bex);
        return f;
    }
    if (family == CGG_BINOMIAL) {
        double f = m[7] * JET_IFACT[7];
#pragma unroll
        for (int k = 6; k >= 1; --k) f = fma(f, h, m[k] * JET_IFACT[k]);
        f = fma(f, h, m[0]);
        // remainder: G_8 S_8 a^8 / 8!;  moment k: rounding <= ce G_k S_k (terms are bounded by |xs|^k G_k)
        double pw = a, bmom = fabs(m[0]);
#pragma unroll
        for (int k = 1; k <= 7; ++k) { bmom = fma(JET_G[k] * JET_IFACT[k] * cst[1 + k], pw, bmom); pw *= a; }
        const double bt = JET_G[8] * JET_IFACT[8] * cst[9] * pw;
        // exact pass: |l'| <= 1, so the rounding of t costs <= eps (|eta| + 2 |x delta|); softplus and sums relative to |f|
        const double bex = JET_EPS * (m[8] + 2.0 * a * cst[2]) + ce * (fabs(f) + bt);
        B = 1.01 * (bt + ce * bmom + bex);
        if (!(a <= JET_AMAX) || m[9] != 0.0) B = INFINITY;
        return f;
    }
    double f = m[6] * JET_IFACT[6];
#pragma unroll
    for (int k = 5; k >= 1; --k) f = fma(f, h, m[k] * JET_IFACT[k]);
    f = fma(f, h, m[0]);
    const double ea = exp(a);
    const double a2 = a * a, a4 = a2 * a2;
    const double bt = ea * (a4 * a2 * a) * JET_IFACT[7] * m[7] * (1.0 + 1e-6);
    // every term of every moment, of the exact pass and of its t-rounding is bounded by e^a (1 + 2a) (|y| + mu)(|eta| + 1)
    const double brnd = (JET_CROUND + 50.0) * JET_EPS * ea * (1.0 + 2.0 * a) * m[8];
    B = 1.01 * (bt + brnd);
    if (!(a <= JET_AMAX) || m[9] != 0.0) B = INFINITY;
    return f;
}

}  // namespace cgg
