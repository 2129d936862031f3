"""SURVEY.md 8(f) rows 2 and 4 on the GPU: the wider family / prior coverage of the reference's vignettes (binomial with
the probit link, negative binomial, gamma / exponential priors, lists of priors: vignettes/pospkg.Rmd:88-108, 132-156,
190-237; R/glm_utils.R:55-57, 113-115), the "naive" linear-predictor mode (R/glm_utils.R:206-208), param_list
(R/mcmcglm.R:183-189, 269, 287) and the runtime-comparison harness (R/measure_performance.R) -- each against the oracle."""
import numpy as np
import pandas as pd
import pytest
import oracle
import mcmcglm_b200 as mg
from helpers import synth
from mcmcglm_b200 import Engine
from mcmcglm_b200.engine import debug_row_terms

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _data(family, n, p, seed):
    if family == "negative_binomial":
        X, _, bt = synth("poisson", n, p, seed=seed)
        rng = np.random.default_rng(seed + 1)
        mu = np.exp(X @ bt)
        y = rng.geometric(1.0 / (1.0 + mu)).astype(np.float64) - 1.0        # dnbinom(size = 1, mu): geometric on 0, 1, ...
        return X, y, bt
    if family == "binomial_probit":
        X, _, bt = synth("binomial", n, p, seed=seed)
        from scipy.special import ndtr
        y = (np.random.default_rng(seed + 1).random(n) < ndtr(X @ bt)).astype(np.float64)
        return X, y, bt
    return synth(family, n, p, seed=seed)


ENGINE_FAMILY = {"negative_binomial": dict(family="negative_binomial"), "binomial_probit": dict(family="binomial", link="probit"),
                 "poisson": dict(family="poisson"), "gaussian": dict(family="gaussian"), "binomial": dict(family="binomial")}
# (engine kwargs, oracle kwargs) of the prior cases
PRIORS = {
    "normal": (dict(prior="normal", prior_mu=0.0, prior_sigma=1.0), dict(prior="normal", prior_mu=0.0, prior_sigma=1.0)),
    "gamma": (dict(prior="gamma", prior_mu=2.0, prior_sigma=1.5), dict(prior="gamma", prior_mu=2.0, prior_sigma=1.5)),
    "exponential": (dict(prior="exponential", prior_sigma=0.7), dict(prior="exponential", prior_sigma=0.7)),
    "list3": (dict(prior="normal", prior_mu=0.1, prior_sigma=2.0, more_priors=(("laplace", 0.0, 1.0, 1.0), ("student_t", 0.0, 1.5, 5.0))),
              dict(prior="normal", prior_mu=0.1, prior_sigma=2.0, more_priors=(("laplace", 0.0, 1.0, 1.0), ("student_t", 0.0, 1.5, 5.0)))),
}


@pytest.mark.parametrize("family", ["negative_binomial", "binomial_probit"])
def test_row_terms_of_the_new_families(family):
    rng = np.random.default_rng(2)
    if family == "negative_binomial":
        eta = np.concatenate([rng.uniform(-8, 6, 3000), [-40.0, -36.5, 30.0, 100.0, 700.0]])
        y = rng.geometric(0.3, eta.size).astype(np.float64) - 1.0
    else:
        eta = np.concatenate([rng.uniform(-9, 9, 3000), [-8.2, 8.2, 0.0, -0.0]])
        y = (rng.random(eta.size) < 0.5).astype(np.float64)
    got = debug_row_terms(family, y, eta)
    ref = oracle.log_density(family, oracle.linkinv(family, eta), y)
    tol = 8 if family == "negative_binomial" else 32      # CUDA normcdf (<= 5 ulp of p) against the oracle's erfc-based pnorm
    assert np.all(np.abs(got - ref) <= tol * EPS * np.maximum(np.abs(ref), 1.0)), np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1))
    if family == "negative_binomial":
        assert debug_row_terms(family, np.array([2.0, 0.0]), np.array([800.0, 800.0])).tolist() == [-np.inf, -np.inf]


@pytest.mark.parametrize("family,prior", [("negative_binomial", "normal"), ("binomial_probit", "list3"), ("poisson", "gamma"),
                                          ("gaussian", "exponential"), ("binomial", "list3"), ("negative_binomial", "gamma")])
def test_g1_and_g2_wider_families_and_priors(family, prior):
    n, p, C, iters = 4001, 4, 2, 20
    X, y, bt = _data(family, n, p, seed=7)
    ekw, okw = PRIORS[prior]
    positive = prior in ("gamma", "exponential")
    m = oracle.make_model(family, sd=1.0, **okw)
    rng = np.random.default_rng(3)
    beta0 = np.abs(0.4 * rng.standard_normal((C, p))) + 0.05 if positive else 0.3 * rng.standard_normal((C, p))
    with Engine(n, p, w=0.4, n_chains=C, K=6, seed=5, sd=1.0, **ENGINE_FAMILY[family], **ekw) as e:
        e.set_data(X, y)
        for c in range(C):
            e.init_chain(c, beta0[c])
        eta = oracle.init_eta(X, beta0[0])
        for j in (0, p - 1):
            cands = beta0[0, j] + np.array([0.0, 0.3, -0.03, 1e-3, 0.7, -2.0])       # (-2.0 leaves the support of gamma / exponential)
            got = e.log_potential(0, j, cands)
            ref = oracle.log_potential(m, X, y, beta0[0], eta, j, cands)
            fin = np.isfinite(ref)
            assert np.array_equal(fin, np.isfinite(got)) and np.array_equal(got[~fin], ref[~fin])        # -Inf outside the support
            assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-12 * np.abs(ref[fin])), (got - ref) / ref
        if positive:
            assert not np.isfinite(ref[-1])
        S, st = e.run(iters)
        for c in range(C):
            r = oracle.run_chain(m, X, y, beta0[c], w=0.4, n_iter=iters, seed=5, chain=c)
            assert r["rc"] == 0
            assert np.max(np.abs(S[c] - r["samples"])) <= 1e-9, (family, prior, c)
            assert st["uniforms_used"][c] == r["uniforms_used"]
            assert e.chain_stats(c)["ref_evals"] == r["n_eval"]
        if positive:
            assert np.all(S > 0.0)               # the sampler never leaves the support


def test_front_door_with_the_vignette_style_models():
    """pospkg.Rmd-style calls: probit, negative binomial (theta ignored like the reference), a gamma prior, a list of priors."""
    X, y, bt = _data("binomial_probit", 3000, 3, seed=11)
    dat = pd.DataFrame({"Y": y, "X1": X[:, 1], "X2": X[:, 2]})
    fit = mg.mcmcglm("Y ~ .", mg.binomial(link="probit"), dat, mg.dist_normal(0, 1), w=0.3, n_samples=300, burnin=50, seed=1)
    est = mg.coef(fit).to_numpy()[0]
    assert np.all(np.abs(est - bt) < 0.25)
    Xn, yn, btn = _data("negative_binomial", 3000, 3, seed=12)
    datn = pd.DataFrame({"Y": yn, "X1": Xn[:, 1], "X2": Xn[:, 2]})
    fitn = mg.mcmcglm("Y ~ .", mg.negative_binomial(3), datn, [mg.dist_normal(0, 1), mg.dist_laplace(0, 1), mg.dist_student_t(4)],
                      w=0.3, n_samples=300, burnin=50, seed=2)
    assert np.all(np.abs(mg.coef(fitn).to_numpy()[0] - btn) < 0.3)
    with pytest.raises(ValueError, match="list length"):
        mg.mcmcglm("Y ~ .", "poisson", datn, [mg.dist_normal(0, 1)] * 2, w=0.3)
    fitg = mg.mcmcglm("Y ~ . - 1", "poisson", pd.DataFrame({"Y": np.random.default_rng(0).poisson(2.0, 500), "X1": np.ones(500)}),
                      mg.dist_gamma(2, 1), w=0.3, n_samples=200, burnin=20, seed=3)
    s = mg.samples(fitg)["X1"].to_numpy()
    assert np.all(s > 0) and abs(np.exp(s[50:]).mean() - 2.0) < 0.3


def test_naive_linear_predictor_mode():
    """linear_predictor_calc = "naive": eta is recomputed as X %*% beta (O(n p)) before every pass over the rows.  Same
    chain as the update mode up to the rounding of eta; the operator form matches the oracle's naive branch."""
    X, y, bt = synth("binomial", 3000, 5, seed=3)
    m = oracle.make_model("binomial", prior="laplace")
    beta0 = 0.2 * np.random.default_rng(1).standard_normal(5)
    ref = oracle.run_chain(m, X, y, beta0, w=0.4, n_iter=15, seed=4, chain=0)
    with Engine(3000, 5, family="binomial", prior="laplace", w=0.4, K=6, seed=4, naive=True) as e:
        e.set_data(X, y)
        e.init_chain(0, beta0)
        S, st = e.run(15)
        beta, eta = e.state(0)
    assert np.max(np.abs(S[0] - ref["samples"])) <= 1e-8 and st["ref_evals"] == ref["n_eval"]
    assert st["launches"] >= 2 * st["passes"]                  # a GEMV launch (plus bookkeeping) for every pass
    assert np.max(np.abs(eta - oracle.init_eta(X, beta))) == 0.0     # eta IS X beta, column order like the oracle's GEMV
    got = mg.log_potential_from_betaj(0.3, 2, beta0, np.zeros(3000), y, X, mg.binomial, mg.dist_laplace(0, 1), "naive")
    want = oracle.log_potential_naive(m, X, y, beta0, 1, 0.3)
    assert abs(got - want) <= 1e-12 * abs(want)


def test_param_list_and_the_comparison_harness():
    X, y, bt = synth("poisson", 500, 3, seed=5)
    dat = pd.DataFrame({"Y": y, "X1": X[:, 1], "X2": X[:, 2]})
    fit = mg.mcmcglm("Y ~ .", "poisson", dat, mg.dist_normal(0, 1), w=0.3, n_samples=12, burnin=4, seed=6, keep_param_list=True)
    pl = fit.param_list
    assert len(pl) == 13 and pl.names[:3] == ["init", "burnin1", "iteration1"] and pl.names[-1] is None      # quirk Q4
    S = mg.samples(fit).iloc[:, :3].to_numpy()
    for k in (0, 5, 12):
        assert np.array_equal(pl[k]["beta"], S[k])
        assert np.max(np.abs(pl[k]["eta"] - X @ S[k])) < 1e-12
        assert np.allclose(pl[k]["mu"], np.maximum(np.exp(pl[k]["eta"]), EPS))
    assert np.array_equal(pl["init"]["beta"], S[0])
    plain = mg.mcmcglm("Y ~ .", "poisson", dat, mg.dist_normal(0, 1), w=0.3, n_samples=12, burnin=4, seed=6)
    assert plain.param_list is None and np.array_equal(mg.samples(plain).iloc[:, :3].to_numpy(), S)
    res = mg.compare_eta_comptime_across_nvars([2, 12], n=100, n_samples=2, burnin=0, rng=np.random.default_rng(0))
    assert list(res["linear_predictor_calc"]) == ["update", "naive"] * 2 and list(res["n_vars"]) == [2, 2, 12, 12]
    assert np.all(res["time"] > 0) and set(res["w"]) == {0.5}
