"""Accuracy of the per-row log-density terms the kernels accumulate, against 40-digit mpmath values of the
reference's formulas (dbinom/dpois/dnorm through the canonical inverse links, incl. R's logit clamps)."""
import numpy as np
import mpmath as mp
import pytest
from mcmcglm_b200.engine import debug_row_terms

pytestmark = pytest.mark.gpu
mp.mp.dps = 40
EPS = np.finfo(float).eps


def _ulp_err(got, exact):
    exact_f = np.array([float(e) for e in exact])
    scale = np.maximum(np.abs(exact_f), 1.0) * EPS
    return np.array([abs(float(mp.mpf(float(g)) - e)) for g, e in zip(got, exact)]) / scale


def test_binomial_row_term_vs_mpmath():
    rng = np.random.default_rng(0)
    eta = np.concatenate([np.linspace(-29.99, 29.99, 1501), rng.uniform(-6, 6, 1500), rng.uniform(-1e-4, 1e-4, 100),
                          [0.0, -0.0, 29.999999, -29.999999]])
    y = (rng.random(eta.size) < 0.5).astype(float)
    got = debug_row_terms("binomial", y, eta)
    exact = [yi * mp.mpf(float(e)) - mp.log1p(mp.exp(mp.mpf(float(e)))) for yi, e in zip(y, eta)]
    err = _ulp_err(got, exact)
    rform = (y == 0.0) & (eta > 8.0)        # R's log(1 - p) regime: checked against the literal form below
    assert err[~rform].max() <= 2.0, err[~rform].max()      # <= 2 ulp of max(|value|, 1)


def test_binomial_row_term_follows_r_log_q_form():
    """y = 0, eta > 8: R computes log(q) with q = 1 - p and p = e / (1 + e) already rounded (stats logit_linkinv +
    nmath dbinom_raw), which differs from the smooth -softplus(eta) by up to 2^-54 (1 + e^eta): 5e-8 at eta = 20.  The
    kernels reproduce R's value: the same double q in all but a few rows near a rounding boundary of p."""
    import oracle
    rng = np.random.default_rng(4)
    eta = np.concatenate([rng.uniform(8.0, 30.0, 4000), np.linspace(8.001, 29.999, 1000)])
    y = np.zeros(eta.size)
    got = debug_row_terms("binomial", y, eta)
    ref = oracle.log_density("binomial", oracle.linkinv("binomial", eta), y)
    smooth = -(eta + np.log1p(np.exp(-eta)))
    assert np.max(np.abs(ref - smooth)[eta > 20]) > 1e-9          # the effect is real ...
    d = np.abs(got - ref)
    assert np.all(d <= 2.0 ** -53 * (1.0 + np.exp(eta)) + 4 * EPS * np.abs(ref))      # ... never off by more than one grid step of p
    assert np.mean(d <= 4 * EPS * np.abs(ref)) >= 0.995            # ... and bit-level agreement in (nearly) every row


def test_binomial_clamps_follow_stats_logit_linkinv():
    # |eta| > 30: R evaluates with exp(eta) clamped to DBL_EPSILON / 1/DBL_EPSILON (stats family.c)
    eta = np.array([30.5, 100.0, 1e300, np.inf, -30.5, -100.0, -np.inf])
    for yv in (0.0, 1.0):
        got = debug_row_terms("binomial", np.full(eta.size, yv), eta)
        e = [mp.mpf(1) / EPS if v > 0 else mp.mpf(EPS) for v in eta]
        exact = [mp.log(x / (1 + x)) if yv == 1.0 else mp.log(1 / (1 + x)) for x in e]
        assert _ulp_err(got, exact).max() <= 2.0
    assert np.isnan(debug_row_terms("binomial", np.array([1.0, 0.0]), np.array([np.nan, np.nan]))).all()


def test_poisson_and_gaussian_row_terms():
    rng = np.random.default_rng(1)
    eta = np.concatenate([rng.uniform(-8, 6, 2000), [-40.0, -36.5, 700.0]])
    y = rng.poisson(3.0, eta.size).astype(float)
    got = debug_row_terms("poisson", y, eta)
    exact = []
    for yi, e in zip(y, eta):
        mu = max(mp.exp(mp.mpf(float(e))), mp.mpf(EPS))     # pmax(exp(eta), .Machine$double.eps)
        exact.append(yi * mp.log(mu) - mu)
    assert _ulp_err(got, exact).max() <= 4.0
    assert debug_row_terms("poisson", np.array([2.0, 0.0]), np.array([800.0, 800.0])).tolist() == [-np.inf, -np.inf]
    yg, eg = rng.standard_normal(500) * 3, rng.standard_normal(500) * 3
    got = debug_row_terms("gaussian", yg, eg, sd=1.7)
    exact = [-(((mp.mpf(float(a)) - mp.mpf(float(b))) / mp.mpf(1.7)) ** 2) / 2 for a, b in zip(yg, eg)]
    assert _ulp_err(got, exact).max() <= 4.0
