/*
 * oracle.c -- CPU restatement of the mcmcglm CGGibbs hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mcmcglm_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED for gaussian + normal prior + slice_stepping_out(w, max=Inf): together
 * with oracle/r_rng.py it reproduces every number the reference prints in its README under seed 42
 * (README.md:73-120: head(samples), coef, quantile) -- see tests/golden/readme_gaussian.json and
 * tests/test_oracle_golden.py.  UNPINNED (no reference output exists; R, qslice, distributional
 * are not installed here) for binomial/poisson likelihoods and laplace/student-t priors: those
 * follow R nmath's published algorithms and are cross-checked against scipy/mpmath only.
 *
 * Every function cites the reference file:line (relative to the mcmcglm repo) or the external
 * dependency it restates.  Arithmetic is deliberately literal: R evaluates `eta + X_j * diff` as
 * two separately rounded vector ops (never an FMA) and `sum()` accumulates in long double, so this
 * file must be compiled with -ffp-contract=off and uses long double accumulators.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_GAUSSIAN 0
#define ORC_BINOMIAL 1        /* binomial, logit link */
#define ORC_POISSON  2
#define ORC_NEGBIN   3        /* MASS::negative.binomial(theta), log link; the reference evaluates dnbinom(size = 1) (R/glm_utils.R:55-57) */
#define ORC_BINOMIAL_PROBIT 4 /* binomial(link = "probit") (vignettes/pospkg.Rmd:88-108) */
#define ORC_NORMAL    0
#define ORC_LAPLACE   1
#define ORC_STUDENT_T 2
#define ORC_GAMMA       3     /* distributional::dist_gamma(shape, rate): a = shape, b = rate */
#define ORC_EXPONENTIAL 4     /* distributional::dist_exponential(rate): b = rate */
#define ORC_MAX_PRIORS 8

#define ORC_OK            0
#define ORC_E_NAN        -1   /* R would stop(): `while (NA)` / `if (NA)` */
#define ORC_E_STREAM     -2   /* replay stream exhausted */
#define ORC_E_NOTERM     -3   /* slice loop did not terminate within the guard */
#define ORC_E_ARG        -4

#define M_LN_SQRT_2PI_ 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#define M_2PI_         6.283185307179586476925286766559
#define ORC_NEG_INF    (-INFINITY)

typedef struct {
    int family;      /* ORC_GAUSSIAN | ORC_BINOMIAL | ORC_POISSON | ORC_NEGBIN | ORC_BINOMIAL_PROBIT */
    double sd;       /* log_likelihood_extra_args$sd (R/mcmcglm.R:151), gaussian only */
    int prior;       /* first (or only) prior: ORC_NORMAL | ORC_LAPLACE | ORC_STUDENT_T | ORC_GAMMA | ORC_EXPONENTIAL */
    double pmu, psigma, pdf;
    /* a LIST of priors (R/glm_utils.R:113-115, quirk Q6): every prior of the list is evaluated at EVERY coordinate and all
     * of it is summed, i.e. the coordinates are iid with density prod_k prior_k.  n_more further components after the first */
    int n_more;
    int more_prior[ORC_MAX_PRIORS - 1];
    double more_a[ORC_MAX_PRIORS - 1], more_b[ORC_MAX_PRIORS - 1], more_c[ORC_MAX_PRIORS - 1];
} orc_model;

/* ------------------------------------------------------------------ nmath pieces */

/* nmath/stirlerr.c: stirlerr(n) = log(n!) - log( sqrt(2*pi*n)*(n/e)^n ).  The sferr_halves table
 * is rebuilt in long double from its definition rather than typed in. */
static double sferr_halves[31];
static int sferr_ready = 0;
static void sferr_init(void) {
    for (int i = 1; i <= 30; ++i) {
        long double n = 0.5L * i;
        long double v = lgammal(n + 1.0L) - (n + 0.5L) * logl(n) + n - 0.5L * logl(2.0L * acosl(-1.0L));
        sferr_halves[i] = (double)v;
    }
    sferr_halves[0] = 0.0;
    sferr_ready = 1;
}

double orc_stirlerr(double n) {
    const double S0 = 0.083333333333333333333;        /* 1/12 */
    const double S1 = 0.00277777777777777777778;      /* 1/360 */
    const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
    const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
    const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
    double nn;
    if (!sferr_ready) sferr_init();
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (int)nn) return sferr_halves[(int)nn];
        return lgamma(n + 1.) - (n + 0.5) * log(n) + n - M_LN_SQRT_2PI_;
    }
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* nmath/bd0.c: bd0(x, np) = x log(x/np) + np - x, evaluated stably near x == np */
double orc_bd0(double x, double np) {
    double ej, s, s1, v;
    int j;
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        v = (x - np) / (x + np);
        s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        ej = 2 * x * v;
        v = v * v;
        for (j = 1; j < 1000; j++) {
            ej *= v;
            s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

/* nmath/dnorm.c (give_log = TRUE) */
double orc_dnorm_log(double x, double mu, double sigma) {
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (sigma < 0) return NAN;
    if (!isfinite(sigma)) return ORC_NEG_INF;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma == 0) return (x == mu) ? INFINITY : ORC_NEG_INF;
    x = (x - mu) / sigma;
    if (!isfinite(x)) return ORC_NEG_INF;
    x = fabs(x);
    if (x >= 2 * sqrt(DBL_MAX)) return ORC_NEG_INF;
    return -(M_LN_SQRT_2PI_ + 0.5 * x * x + log(sigma));
}

static int r_nonint(double x) { return fabs(x - nearbyint(x)) > 1e-7 * fmax(1., fabs(x)); }

/* nmath/dbinom.c: dbinom_raw(x, n, p, q, give_log = TRUE) */
static double dbinom_raw_log(double x, double n, double p, double q) {
    double lf, lc;
    if (p == 0) return (x == 0) ? 0. : ORC_NEG_INF;
    if (q == 0) return (x == n) ? 0. : ORC_NEG_INF;
    if (x == 0) {
        if (n == 0) return 0.;
        lc = (p < 0.1) ? -orc_bd0(n, n * q) - n * p : n * log(q);
        return lc;
    }
    if (x == n) {
        lc = (q < 0.1) ? -orc_bd0(n, n * p) - n * q : n * log(p);
        return lc;
    }
    if (x < 0 || x > n) return ORC_NEG_INF;
    lc = orc_stirlerr(n) - orc_stirlerr(x) - orc_stirlerr(n - x) - orc_bd0(x, n * p) - orc_bd0(n - x, n * q);
    lf = log(M_2PI_) + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

/* nmath/dbinom.c: dbinom(x, n, p, give_log = TRUE) */
double orc_dbinom_log(double x, double n, double p) {
    if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
    if (p < 0 || p > 1 || n < 0 || r_nonint(n)) return NAN;
    if (r_nonint(x)) return ORC_NEG_INF; /* R warns and returns R_D__0 */
    if (x < 0 || !isfinite(x)) return ORC_NEG_INF;
    n = nearbyint(n);
    x = nearbyint(x);
    return dbinom_raw_log(x, n, p, 1 - p);
}

/* nmath/dpois.c: dpois_raw(x, lambda, give_log = TRUE).  This is the R <= 4.0.x body
 * (-stirlerr(x) - bd0(x, lambda)); R >= 4.1 computes the same quantity through ebd0(), a
 * split-precision table variant that differs by <= ~1 ulp. */
static double dpois_raw_log(double x, double lambda) {
    if (lambda == 0) return (x == 0) ? 0. : ORC_NEG_INF;
    if (!isfinite(lambda)) return ORC_NEG_INF;
    if (x < 0) return ORC_NEG_INF;
    if (x <= lambda * DBL_MIN) return -lambda;
    if (lambda < x * DBL_MIN) {
        if (!isfinite(x)) return ORC_NEG_INF;
        return -lambda + x * log(lambda) - lgamma(x + 1);
    }
    return -0.5 * log(M_2PI_ * x) + (-orc_stirlerr(x) - orc_bd0(x, lambda));
}

/* nmath/dpois.c: dpois(x, lambda, give_log = TRUE) */
double orc_dpois_log(double x, double lambda) {
    if (isnan(x) || isnan(lambda)) return x + lambda;
    if (lambda < 0) return NAN;
    if (r_nonint(x)) return ORC_NEG_INF;
    if (x < 0 || !isfinite(x)) return ORC_NEG_INF;
    x = nearbyint(x);
    return dpois_raw_log(x, lambda);
}

/* nmath/dt.c: dt(x, n, give_log = TRUE) */
double orc_dt_log(double x, double n) {
    if (isnan(x) || isnan(n)) return x + n;
    if (n <= 0) return NAN;
    if (!isfinite(x)) return ORC_NEG_INF;
    if (!isfinite(n)) return orc_dnorm_log(x, 0., 1.);
    double u;
    double t = -orc_bd0(n / 2., (n + 1) / 2.) + orc_stirlerr((n + 1) / 2.) - orc_stirlerr(n / 2.);
    double x2n = x * x / n, ax = 0., l_x2n;
    int lrg_x2n = (x2n > 1. / DBL_EPSILON);
    if (lrg_x2n) {
        ax = fabs(x);
        l_x2n = log(ax) - log(n) / 2.;
        u = n * l_x2n;
    } else if (x2n > 0.2) {
        l_x2n = log(1 + x2n) / 2.;
        u = n * l_x2n;
    } else {
        l_x2n = log1p(x2n) / 2.;
        u = -orc_bd0(n / 2., (n + x * x) / 2.) + x * x / 2.;
    }
    (void)ax;
    return t - u - (M_LN_SQRT_2PI_ + l_x2n);
}

/* nmath/dnbinom.c: dnbinom_mu(x, size, mu, give_log = TRUE) */
double orc_dnbinom_mu_log(double x, double size, double mu) {
    if (isnan(x) || isnan(size) || isnan(mu)) return x + size + mu;
    if (mu < 0 || size < 0) return NAN;
    if (r_nonint(x)) return ORC_NEG_INF;
    if (x < 0 || !isfinite(x)) return ORC_NEG_INF;
    if (x == 0 && size == 0) return 0.;
    x = nearbyint(x);
    if (!isfinite(size)) return orc_dpois_log(x, mu);
    if (x == 0) return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    if (x < 1e-10 * size) {
        double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
        return x * p - mu - lgamma(x + 1) + log1p(x * (x - 1) / (2 * size));
    } else {
        double p = size / (size + x), ans = dbinom_raw_log(size, x + size, size / (size + mu), mu / (size + mu));
        return log(p) + ans;
    }
}

/* nmath/pnorm.c, lower tail: restated through the C library's erfc (nmath uses Cody's rational approximations; the two
 * agree to ~1e-16 relative over the range stats' probit link allows, |x| <= 8.13) */
double orc_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* nmath/dgamma.c: dgamma(x, shape, scale, give_log = TRUE) */
double orc_dgamma_log(double x, double shape, double scale) {
    double pr;
    if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
    if (shape < 0 || scale <= 0) return NAN;
    if (x < 0) return ORC_NEG_INF;
    if (shape == 0) return (x == 0) ? INFINITY : ORC_NEG_INF;
    if (x == 0) {
        if (shape < 1) return INFINITY;
        if (shape > 1) return ORC_NEG_INF;
        return -log(scale);
    }
    if (shape < 1) {
        pr = dpois_raw_log(shape, x / scale);
        return pr + (isfinite(shape / x) ? log(shape / x) : log(shape) - log(x));
    }
    pr = dpois_raw_log(shape - 1, x / scale);
    return pr - log(scale);
}

/* nmath/dexp.c: dexp(x, scale, give_log = TRUE) */
double orc_dexp_log(double x, double scale) {
    if (isnan(x) || isnan(scale)) return x + scale;
    if (scale <= 0.0) return NAN;
    if (x < 0.) return ORC_NEG_INF;
    return (-x / scale) - log(scale);
}

/* ------------------------------------------------------------------ links (stats) */

/* family$linkinv, R/glm_utils.R:210.  gaussian(): identity.  binomial(): stats/src/family.c
 * logit_linkinv (clamps at |eta| > 30, x/(1+x)).  poisson(): pmax(exp(eta), .Machine$double.eps) */
void orc_linkinv(int family, int64_t n, const double *eta, double *mu) {
    const double THRESH = 30., MTHRESH = -30., INVEPS = 1 / DBL_EPSILON;
    int64_t i;
    switch (family) {
    case ORC_GAUSSIAN:
        for (i = 0; i < n; ++i) mu[i] = eta[i];
        break;
    case ORC_BINOMIAL:
        for (i = 0; i < n; ++i) {
            double etai = eta[i];
            double tmp = (etai < MTHRESH) ? DBL_EPSILON : ((etai > THRESH) ? INVEPS : exp(etai));
            mu[i] = tmp / (1 + tmp);
        }
        break;
    case ORC_BINOMIAL_PROBIT: {
        /* stats::binomial(link = "probit")$linkinv (make.link): thresh <- -qnorm(.Machine$double.eps);
         * eta <- pmin(pmax(eta, -thresh), thresh); pnorm(eta) */
        const double thresh = 8.125890664701906;
        for (i = 0; i < n; ++i) {
            double e = eta[i];
            if (!isnan(e)) e = (e < -thresh) ? -thresh : ((e > thresh) ? thresh : e);
            mu[i] = orc_pnorm(e);
        }
        break;
    }
    default:   /* poisson()$linkinv and MASS::negative.binomial()$linkinv (log link): pmax(exp(eta), .Machine$double.eps) */
        for (i = 0; i < n; ++i) {
            double e = exp(eta[i]);
            mu[i] = (e > DBL_EPSILON) ? e : DBL_EPSILON; /* pmax keeps NaN; exp never returns it for finite eta */
            if (isnan(e)) mu[i] = e;
        }
    }
}

/* log_density.<family>, R/glm_utils.R:40-52 */
double orc_log_density(int family, double mu, double y, double sd) {
    switch (family) {
    case ORC_GAUSSIAN: return orc_dnorm_log(y, mu, sd);
    case ORC_BINOMIAL: case ORC_BINOMIAL_PROBIT: return orc_dbinom_log(y, 1.0, mu);
    case ORC_NEGBIN:   return orc_dnbinom_mu_log(y, 1.0, mu);      /* R/glm_utils.R:55-57: size = 1 whatever theta is */
    default:           return orc_dpois_log(y, mu);
    }
}

/* log_likelihood, R/glm_utils.R:93-99: sum(log_density(...)); R's sum() accumulates in LDOUBLE */
double orc_log_likelihood(int family, int64_t n, const double *mu, const double *y, double sd) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += orc_log_density(family, mu[i], y[i], sd);
    return (double)s;
}

/* distributional::density(<dist>, at, log = TRUE) for one coordinate (external, unpinned):
 * normal -> dnorm(at, mu, sigma, log=TRUE); laplace -> -log(2 sigma) - |at - mu| / sigma;
 * student_t(df, mu, sigma) -> dt((at - mu)/sigma, df, log=TRUE) - log(sigma); gamma(shape, rate) -> dgamma(at, shape, rate, log=TRUE);
 * exponential(rate) -> dexp(at, rate, log=TRUE).  A list of priors sums its components at every coordinate (quirk Q6). */
static double prior_component(int kind, double a, double b, double c, double at) {
    switch (kind) {
    case ORC_NORMAL:      return orc_dnorm_log(at, a, b);
    case ORC_LAPLACE:     return -log(2 * b) - fabs(at - a) / b;
    case ORC_STUDENT_T:   return orc_dt_log((at - a) / b, c) - log(b);
    case ORC_GAMMA:       return orc_dgamma_log(at, a, 1.0 / b);      /* distributional: dgamma(x, shape, rate) */
    default:              return orc_dexp_log(at, 1.0 / b);           /* distributional: dexp(x, rate) */
    }
}
double orc_prior_log_density1(const orc_model *m, double at) {
    double v = prior_component(m->prior, m->pmu, m->psigma, m->pdf, at);
    for (int k = 0; k < m->n_more; ++k) v += prior_component(m->more_prior[k], m->more_a[k], m->more_b[k], m->more_c[k], at);
    return v;
}

/* log_prior_density.default, R/glm_utils.R:108-110: the prior is evaluated at ALL p coordinates
 * on every call (quirk Q5) and summed with sum() */
double orc_log_prior_density(const orc_model *m, int64_t p, const double *beta) {
    long double s = 0.0L;
    for (int64_t l = 0; l < p; ++l) s += orc_prior_log_density1(m, beta[l]);
    return (double)s;
}

/* update_linear_predictor, R/glm_utils.R:126-132: two vectorised ops => two roundings, no FMA */
void orc_update_linear_predictor(int64_t n, double new_beta_j, double current_beta_j,
                                 const double *current_eta, const double *X_j, double *new_eta) {
    double diff_beta = new_beta_j - current_beta_j;
    for (int64_t i = 0; i < n; ++i) {
        volatile double prod = X_j[i] * diff_beta;
        new_eta[i] = current_eta[i] + prod;
    }
}

/* log_potential_from_betaj, R/glm_utils.R:187-218, linear_predictor_calc = "update".
 * scratch: 2*n + p doubles.  j is 0-based. */
double orc_log_potential(const orc_model *m, int64_t n, int64_t p, const double *X, int64_t ldx,
                         const double *y, const double *beta, const double *eta, int64_t j,
                         double new_beta_j, double *scratch) {
    double *new_eta = scratch, *new_mu = scratch + n, *new_beta = scratch + 2 * n;
    memcpy(new_beta, beta, (size_t)p * sizeof(double));
    new_beta[j] = new_beta_j;                                                 /* :197-198 */
    orc_update_linear_predictor(n, new_beta_j, beta[j], eta, X + j * ldx, new_eta); /* :200-205 */
    orc_linkinv(m->family, n, new_eta, new_mu);                               /* :210 */
    double ll = orc_log_likelihood(m->family, n, new_mu, y, m->sd);          /* :212 */
    double lp = orc_log_prior_density(m, p, new_beta);                        /* :214-215 */
    return ll + lp;                                                           /* :217 */
}

/* "naive" branch, R/glm_utils.R:206-208: new_eta <- X %*% new_beta (reference BLAS dgemv; plain
 * column-ordered accumulation here) */
double orc_log_potential_naive(const orc_model *m, int64_t n, int64_t p, const double *X, int64_t ldx,
                               const double *y, const double *beta, int64_t j, double new_beta_j,
                               double *scratch) {
    double *new_eta = scratch, *new_mu = scratch + n, *new_beta = scratch + 2 * n;
    memcpy(new_beta, beta, (size_t)p * sizeof(double));
    new_beta[j] = new_beta_j;
    for (int64_t i = 0; i < n; ++i) new_eta[i] = 0.0;
    for (int64_t l = 0; l < p; ++l)
        for (int64_t i = 0; i < n; ++i) new_eta[i] += X[l * ldx + i] * new_beta[l];
    orc_linkinv(m->family, n, new_eta, new_mu);
    return orc_log_likelihood(m->family, n, new_mu, y, m->sd) + orc_log_prior_density(m, p, new_beta);
}

/* init_eta <- drop(X %*% init_beta), R/mcmcglm.R:215 */
void orc_init_eta(int64_t n, int64_t p, const double *X, int64_t ldx, const double *beta, double *eta) {
    for (int64_t i = 0; i < n; ++i) eta[i] = 0.0;
    for (int64_t l = 0; l < p; ++l)
        for (int64_t i = 0; i < n; ++i) eta[i] += X[l * ldx + i] * beta[l];
}

/* ------------------------------------------------------------------ uniform streams */

/* Philox4x32-10 (Salmon et al. 2011).  Shared definition with the device code: uniform #idx of
 * chain c under seed s is u = ((x >> 12) + 0.5) * 2^-52 with x = (r0 << 32 | r1) of
 * philox4x32_10(ctr = {idx_lo, idx_hi, c, 0x43474742}, key = {s_lo, s_hi}). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_philox_uniform(uint64_t seed, uint32_t chain, uint64_t idx) {
    uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), chain, 0x43474742u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
    orc_philox4x32_10(ctr, key, r);
    uint64_t x = ((uint64_t)r[0] << 32) | r[1];
    return ((double)(x >> 12) + 0.5) * 0x1p-52;
}

typedef struct {
    const double *u;   /* replay buffer (R's runif draws, or any recorded stream); NULL => Philox */
    uint64_t n, pos;   /* pos = number of uniforms consumed so far */
    uint64_t seed;
    uint32_t chain;
} orc_stream;

static int stream_next(orc_stream *s, double *out) {
    if (s->u) {
        if (s->pos >= s->n) return ORC_E_STREAM;
        *out = s->u[s->pos++];
    } else {
        *out = orc_philox_uniform(s->seed, s->chain, s->pos++);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ slice sampler (qslice) */

typedef double (*orc_target)(double x, void *ctx);

typedef struct {
    int64_t n_eval;      /* qslice's nEvaluations */
    int64_t n_stepout;   /* bracket expansions (L <- L - w or R <- R + w) */
    int64_t n_shrink;    /* shrink proposals, accepted one included */
} orc_slice_stats;

/* qslice::slice_stepping_out(x, log_target, w, max = Inf) -- EXTERNAL (CRAN qslice, unpinned in
 * DESCRIPTION:12-23), call site R/mcmcglm.R:258-261.  Neal (2003) Fig. 3 + Fig. 5 on the log
 * scale.  Draw order: slice level, bracket offset, [J when max is finite], one per shrink step.
 * max_steps < 0 means max = Inf.  Pinned by the README golden run (tests/golden). */
int orc_slice_stepping_out(double x, orc_target f, void *ctx, double w, int64_t max_steps,
                           orc_stream *st, double *x_out, double *fx_out, orc_slice_stats *stats) {
    double u, fx, y, L, R, x1, f1;
    int rc;
    int64_t guard = 0;
    const int64_t GUARD = 100000;
    fx = f(x, ctx); stats->n_eval++;
    if (isnan(fx)) return ORC_E_NAN;
    if ((rc = stream_next(st, &u))) return rc;
    y = log(u) + fx;
    if ((rc = stream_next(st, &u))) return rc;
    L = x - u * w;
    R = L + w;
    if (max_steps < 0) {
        for (;;) {
            f1 = f(L, ctx); stats->n_eval++;
            if (isnan(f1)) return ORC_E_NAN;
            if (!(y < f1)) break;
            L = L - w; stats->n_stepout++;
            if (++guard > GUARD) return ORC_E_NOTERM;
        }
        for (;;) {
            f1 = f(R, ctx); stats->n_eval++;
            if (isnan(f1)) return ORC_E_NAN;
            if (!(y < f1)) break;
            R = R + w; stats->n_stepout++;
            if (++guard > GUARD) return ORC_E_NOTERM;
        }
    } else if (max_steps > 0) {
        if ((rc = stream_next(st, &u))) return rc;
        double J = floor(u * (double)max_steps);
        double K = (double)max_steps - 1 - J;
        while (J > 0) {
            f1 = f(L, ctx); stats->n_eval++;
            if (isnan(f1)) return ORC_E_NAN;
            if (!(y < f1)) break;
            L = L - w; J = J - 1; stats->n_stepout++;
        }
        while (K > 0) {
            f1 = f(R, ctx); stats->n_eval++;
            if (isnan(f1)) return ORC_E_NAN;
            if (!(y < f1)) break;
            R = R + w; K = K - 1; stats->n_stepout++;
        }
    }
    for (;;) {
        if ((rc = stream_next(st, &u))) return rc;
        x1 = L + u * (R - L);
        f1 = f(x1, ctx); stats->n_eval++; stats->n_shrink++;
        if (isnan(f1)) return ORC_E_NAN;
        if (y < f1) { *x_out = x1; *fx_out = f1; return ORC_OK; }
        if (x1 < x) L = x1; else R = x1;
        if (++guard > GUARD) return ORC_E_NOTERM;
    }
}

/* ------------------------------------------------------------------ Gibbs sweep */

typedef struct {
    const orc_model *m;
    int64_t n, p, ldx, j;
    const double *X, *y, *beta, *eta;
    double *scratch;
} lp_ctx;

static double lp_target(double b, void *vctx) {
    lp_ctx *c = (lp_ctx *)vctx;
    return orc_log_potential(c->m, c->n, c->p, c->X, c->ldx, c->y, c->beta, c->eta, c->j, b, c->scratch);
}

/* The (k, j) double loop, R/mcmcglm.R:226-274, sample_method = "slice_sampling",
 * linear_predictor_calc = "update".  beta/eta are updated in place; samples is n_iter x p
 * row-major (row = iteration k = 1..n_iter; the caller stores beta0 as row 0 itself).
 * max_updates > 0 stops after that many coordinate updates (used by the bounded CPU baseline).
 * uniforms_used/stats are cumulative outputs.  The dead `mu <- linkinv(eta)` of :269 is computed
 * when compute_mu != 0 so that a timed run does the same work as the reference. */
int orc_run_chain(const orc_model *m, int64_t n, int64_t p, const double *X, int64_t ldx,
                  const double *y, double *beta, double *eta, double w, int64_t max_steps,
                  int64_t n_iter, int64_t max_updates, const double *replay_u, uint64_t n_u,
                  uint64_t seed, uint32_t chain, uint64_t stream_pos0, int compute_mu,
                  double *samples, uint64_t *uniforms_used, orc_slice_stats *stats) {
    if (n <= 0 || p <= 0 || !(w > 0)) return ORC_E_ARG;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(3 * n + p));
    if (!scratch) return ORC_E_ARG;
    double *mu = scratch + 2 * n + p;
    orc_stream st = {replay_u, n_u, stream_pos0, seed, chain};
    lp_ctx ctx = {m, n, p, ldx, 0, X, y, beta, eta, scratch};
    int rc = ORC_OK;
    int64_t done = 0;
    for (int64_t k = 0; k < n_iter && rc == ORC_OK; ++k) {
        for (int64_t j = 0; j < p; ++j) {
            double x1 = 0, f1 = 0;
            ctx.j = j;
            rc = orc_slice_stepping_out(beta[j], lp_target, &ctx, w, max_steps, &st, &x1, &f1, stats);
            if (rc != ORC_OK) break;
            double old = beta[j];
            beta[j] = x1;                                                       /* :264 */
            orc_update_linear_predictor(n, x1, old, eta, X + j * ldx, eta);     /* :265-268 */
            if (compute_mu) orc_linkinv(m->family, n, eta, mu);                 /* :269 */
            if (samples) samples[k * p + j] = x1;                               /* :271 */
            if (max_updates > 0 && ++done >= max_updates) goto out;
        }
    }
out:
    if (uniforms_used) *uniforms_used = st.pos;
    free(scratch);
    return rc;
}

/* K candidates of one coordinate at once: the oracle for cgg_log_potential (parity gate G1) */
int orc_log_potential_batch(const orc_model *m, int64_t n, int64_t p, const double *X, int64_t ldx,
                            const double *y, const double *beta, const double *eta, int64_t j,
                            int K, const double *cand, double *out) {
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(2 * n + p));
    if (!scratch) return ORC_E_ARG;
    for (int k = 0; k < K; ++k)
        out[k] = orc_log_potential(m, n, p, X, ldx, y, beta, eta, j, cand[k], scratch);
    free(scratch);
    return ORC_OK;
}
