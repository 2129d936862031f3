"""Cross-checks of the oracle's restated nmath / distributional / qslice pieces against independent
implementations (scipy, mpmath) and closed forms.  These bound the UNPINNED parts of the oracle."""
import numpy as np
import mpmath as mp
import pytest
from scipy import stats
import oracle
from helpers import synth, PRIOR_CASES

mp.mp.dps = 40


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_uniform_open_interval():
    u = np.array([oracle.philox_uniform(7, 3, i) for i in range(20000)])
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005


def test_stirlerr_and_bd0():
    L = oracle.lib()
    for n in [0.5, 1, 2.5, 7, 15, 15.5, 20, 36, 81, 501, 1e4]:
        exact = mp.loggamma(n + 1) - (n + mp.mpf(0.5)) * mp.log(n) + n - mp.log(mp.sqrt(2 * mp.pi))
        assert abs(L.orc_stirlerr(n) - float(exact)) < 2e-16 * max(1.0, 1.0 / n) + 1e-17
    for x, npv in [(1.0, 0.95), (1.0, 0.5), (400.0, 380.0), (3.0, 2.5), (2.0, 2.5)]:
        exact = x * mp.log(mp.mpf(x) / npv) + npv - x
        assert abs(L.orc_bd0(x, npv) - float(exact)) <= 4e-16 * max(1.0, abs(float(exact)))


def test_densities_vs_scipy():
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for y, lam in zip(rng.poisson(30, 200), rng.gamma(30, 1, 200)):
        assert np.isclose(L.orc_dpois_log(float(y), lam), stats.poisson.logpmf(y, lam), rtol=1e-13, atol=1e-13)
    for y, p in zip(rng.integers(0, 2, 200), rng.random(200)):
        assert np.isclose(L.orc_dbinom_log(float(y), 1.0, p), stats.binom.logpmf(y, 1, p), rtol=1e-14, atol=1e-15)
    for x in rng.standard_normal(100) * 5:
        assert np.isclose(L.orc_dnorm_log(x, 0.3, 1.7), stats.norm.logpdf(x, 0.3, 1.7), rtol=1e-15)
        for df in (1.0, 4.0, 3.5, 30.0):
            assert np.isclose(L.orc_dt_log(x, df), stats.t.logpdf(x, df), rtol=2e-14)
    # R semantics: non-integer / negative counts give log-density -Inf
    assert L.orc_dpois_log(1.5, 2.0) == -np.inf and L.orc_dbinom_log(2.0, 1.0, 0.5) == -np.inf
    assert L.orc_dpois_log(3.0, np.inf) == -np.inf


def test_logit_link_clamps():
    # stats family.c: eta < -30 -> DBL_EPSILON, eta > 30 -> 1/DBL_EPSILON, then x/(1+x)
    eps = np.finfo(float).eps
    mu = oracle.linkinv("binomial", np.array([-31.0, 31.0, -40.0, 400.0, 0.0, 2.0]))
    assert mu[0] == eps / (1 + eps) and mu[2] == mu[0]
    assert mu[1] == (1 / eps) / (1 + 1 / eps) and mu[3] == mu[1]
    assert mu[4] == 0.5 and np.isclose(mu[5], 1 / (1 + np.exp(-2.0)), rtol=1e-16)
    # poisson(): pmax(exp(eta), eps)
    mu = oracle.linkinv("poisson", np.array([-50.0, 0.0, 3.0]))
    assert mu[0] == eps and mu[1] == 1.0 and mu[2] == np.exp(3.0)


def _mp_log_potential(family, prior, X, y, beta, eta, j, b, sd=1.0):
    """High-precision f(b): exact math of R/glm_utils.R:187-218 (no clamps: only used at moderate eta)."""
    nb = [mp.mpf(float(v)) for v in beta]
    nb[j] = mp.mpf(float(b))
    d = mp.mpf(float(b)) - mp.mpf(float(beta[j]))
    tot = mp.mpf(0)
    for i in range(len(y)):
        e = mp.mpf(float(eta[i])) + mp.mpf(float(X[i, j])) * d
        yi = mp.mpf(float(y[i]))
        if family == "gaussian":
            z = (yi - e) / sd
            tot += -(mp.log(mp.sqrt(2 * mp.pi)) + z * z / 2 + mp.log(sd))
        elif family == "binomial":
            tot += yi * e - mp.log1p(mp.exp(e))
        else:
            tot += yi * e - mp.exp(e) - mp.loggamma(yi + 1)
    for v in nb:
        if prior["prior"] == "normal":
            z = (v - prior["prior_mu"]) / prior["prior_sigma"]
            tot += -(mp.log(mp.sqrt(2 * mp.pi)) + z * z / 2 + mp.log(prior["prior_sigma"]))
        elif prior["prior"] == "laplace":
            tot += -mp.log(2 * prior["prior_sigma"]) - abs(v - prior["prior_mu"]) / prior["prior_sigma"]
        else:
            nu = mp.mpf(prior["prior_df"])
            z = (v - prior["prior_mu"]) / prior["prior_sigma"]
            tot += (mp.loggamma((nu + 1) / 2) - mp.loggamma(nu / 2) - mp.log(nu * mp.pi) / 2
                    - (nu + 1) / 2 * mp.log1p(z * z / nu) - mp.log(prior["prior_sigma"]))
    return tot


@pytest.mark.parametrize("family", ["gaussian", "binomial", "poisson"])
@pytest.mark.parametrize("prior", ["normal", "laplace", "student_t"])
def test_log_potential_vs_mpmath(family, prior):
    X, y, bt = synth(family, 400, 5, seed=3)
    rng = np.random.default_rng(5)
    beta = bt + 0.3 * rng.standard_normal(5)
    eta = oracle.init_eta(X, beta)
    m = oracle.make_model(family, sd=1.3, **PRIOR_CASES[prior])
    for j in range(5):
        cands = beta[j] + np.array([-0.5, -0.01, 0.0, 0.2])
        got = oracle.log_potential(m, X, y, beta, eta, j, cands)
        for b, g in zip(cands, got):
            ex = _mp_log_potential(family, PRIOR_CASES[prior], X, y, beta, eta, j, b, sd=1.3)
            assert abs(g - float(ex)) <= 1e-13 * abs(float(ex))


def test_update_equals_naive():
    # linear_predictor_calc "update" and "naive" agree up to rounding (R/glm_utils.R:200-208)
    X, y, bt = synth("binomial", 300, 6, seed=2)
    m = oracle.make_model("binomial", **PRIOR_CASES["normal"])
    eta = oracle.init_eta(X, bt)
    for j in range(6):
        a = oracle.log_potential(m, X, y, bt, eta, j, [bt[j] + 0.37])[0]
        b = oracle.log_potential_naive(m, X, y, bt, j, bt[j] + 0.37)
        assert np.isclose(a, b, rtol=1e-13)


def test_update_linear_predictor_is_two_roundings():
    rng = np.random.default_rng(0)
    eta, xj = rng.standard_normal(1000), rng.standard_normal(1000)
    out = oracle.update_linear_predictor(0.7, 0.1, eta, xj)
    assert np.array_equal(out, eta + xj * (0.7 - 0.1))   # numpy does mul then add, like R


def test_gaussian_posterior_closed_form():
    # known answer from the reference's own formula R/sampling.R:8-9:
    # Sigma_post = (X'X/s^2 + Sigma0^-1)^-1, mu_post = Sigma_post X'Y/s^2  (prior mean 0)
    X, y, _ = synth("gaussian", 500, 3, seed=11)
    m = oracle.make_model("gaussian", sd=1.0, **PRIOR_CASES["normal"])
    cov = np.linalg.inv(X.T @ X + np.eye(3))
    mu = cov @ X.T @ y
    out = oracle.run_chain(m, X, y, np.zeros(3), w=0.5, n_iter=6000, seed=123)
    S = out["samples"][500:]
    se = np.sqrt(np.diag(cov))
    assert np.all(np.abs(S.mean(0) - mu) < 0.15 * se)
    assert np.all(np.abs(S.std(0) / se - 1) < 0.1)


def test_finite_max_draws_extra_uniform():
    X, y, _ = synth("poisson", 200, 3, seed=4)
    m = oracle.make_model("poisson", **PRIOR_CASES["student_t"])
    a = oracle.run_chain(m, X, y, np.zeros(3), w=0.05, n_iter=20, seed=9, max_steps=-1)
    b = oracle.run_chain(m, X, y, np.zeros(3), w=0.05, n_iter=20, seed=9, max_steps=4)
    assert a["rc"] == 0 and b["rc"] == 0
    assert a["uniforms_used"] == 2 * 60 + a["n_shrink"]
    assert b["uniforms_used"] == 3 * 60 + b["n_shrink"]


def test_replay_stream_exhaustion_is_an_error():
    X, y, _ = synth("gaussian", 50, 2, seed=4)
    m = oracle.make_model("gaussian", **PRIOR_CASES["normal"])
    out = oracle.run_chain(m, X, y, np.zeros(2), w=0.5, n_iter=10, replay_u=np.full(5, 0.5))
    assert out["rc"] == oracle.E_STREAM


def test_r_log_q_shortcut_matches_the_literal_form():
    """The device evaluates R's `log(q), q = 1 - p, p = fl(e / (1 + e))` (y = 0, eta in (8, 30]) as
    -softplus(eta) + log1p(rho) with q_R = 1 - fl(1 - T / (1 + T)), T = exp(-eta) (cgg_math.cuh: rform_log1p_rho).
    This is the same arithmetic in numpy against the oracle's literal dbinom form: the same q in all but a few rows,
    and sums that agree to ~1e-17 relative (the smooth -softplus alone is off by 1e-9 at sd(eta) ~ 15)."""
    rng = np.random.default_rng(0)
    eta = rng.uniform(8.0, 30.0, 200_000)
    y = np.zeros(eta.size)
    lit = oracle.log_density("binomial", oracle.linkinv("binomial", eta), y)
    T = np.exp(-eta)
    u = 1 - T * (1 - T * (1 - T * (1 - T)))
    qh = T * u
    qR = 1 - (1 - qh)
    rho = (qR - qh) * (np.exp(eta) * (1 + T))
    mine = -(eta + np.log1p(T)) + rho * (1 - rho * (0.5 - rho / 3))
    q_lit = 1 - oracle.linkinv("binomial", eta)
    assert np.mean(q_lit == qR) > 0.999
    assert np.all(np.abs(mine - lit) <= 2.0 ** -53 * (1 + np.exp(eta)) + 1e-14)
    assert abs(mine.sum() - lit.sum()) <= 1e-15 * abs(lit.sum())
    assert abs(-(eta + np.log1p(T)).sum() - lit.sum()) > 1e-9 * abs(lit.sum())
