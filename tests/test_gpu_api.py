"""The reference-facing interface end to end on the GPU: the README example (reference README.md:62-120)
through mcmcglm()/samples()/coef()/quantile(), the exported operators, and posterior checks (gate G3)."""
import json
import os
import numpy as np
import pandas as pd
import pytest
import oracle
import mcmcglm_b200 as mg
from helpers import synth, PRIOR_CASES

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_readme_example_through_the_front_door():
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    readme = json.load(open(os.path.join(G, "readme_gaussian.json")))
    dat = pd.DataFrame({"Y": z["y"], "X1": z["X"][:, 1], "X2": z["X"][:, 2]})
    # same call as README.md:62-66; the prior draw and R's runif stream are supplied so the chain is R's
    norm = mg.mcmcglm(formula="Y ~ .", family="gaussian", data=dat, beta_prior=mg.dist_normal(0, 1), w=0.5,
                      beta_init=z["beta0"], replay_uniforms=z["uniforms"])
    S = mg.samples(norm)
    assert list(S.columns) == ["(Intercept)", "X1", "X2", "iteration", "burnin"] and len(S) == 501
    assert S["burnin"].sum() == 102                                                     # quirk Q1
    for row, prow in zip(S.iloc[:6, :3].to_numpy(), readme["head_samples"]):            # README.md:114-120
        assert np.allclose(row, prow, rtol=2e-7, atol=1e-9)
    assert np.allclose(mg.coef(norm).to_numpy()[0], list(readme["coef"].values()), rtol=1e-6)   # README.md:93-94
    q = mg.quantile(norm)                                                                # README.md:103-106
    for _, r in q.iterrows():
        ref = readme["quantile"][r["var"]]
        assert np.allclose([r["mean"], r["q_0025"], r["q_05"], r["q_0975"]],
                           [ref["mean"], ref["q_0025"], ref["q_05"], ref["q_0975"]], rtol=1e-6)
    assert norm.w == 0.5 and norm.stats["ref_evals"] == int(z["n_eval"])
    assert "Average of parameter samples" in repr(norm)


@pytest.mark.parametrize("family,prior", [("gaussian", "normal"), ("binomial", "laplace"), ("poisson", "student_t")])
def test_exported_operators(family, prior):
    X, y, bt = synth(family, 4001, 6, seed=5)
    eta = X @ bt
    fam = {"gaussian": mg.gaussian, "binomial": mg.binomial, "poisson": mg.poisson}[family]
    pr = {"normal": mg.dist_normal(0, 1), "laplace": mg.dist_laplace(0, 1), "student_t": mg.dist_student_t(4, 0, 1)}[prior]
    m = oracle.make_model(family, sd=1.0, **PRIOR_CASES[prior])
    for j in (1, 4, 6):                                             # 1-based like the reference
        got = mg.log_potential_from_betaj(bt[j - 1] + 0.1, j, bt, eta, y, X, fam, pr, sd=1.0)
        ref = oracle.log_potential(m, X, y, bt, eta, j - 1, [bt[j - 1] + 0.1])[0]
        assert abs(got - ref) <= 1e-12 * abs(ref)
    new_eta = mg.update_linear_predictor(0.7, bt[2], eta, X[:, 2])
    assert np.array_equal(new_eta, oracle.update_linear_predictor(0.7, bt[2], eta, X[:, 2]))


def test_g3_posterior_matches_closed_form_gaussian():
    # gaussian + normal(0,1) prior: posterior N(mu_post, Sigma_post) with the reference's own formula R/sampling.R:8-9
    X, y, _ = synth("gaussian", 2000, 4, seed=3)
    dat = pd.DataFrame({"Y": y, "A": X[:, 1], "B": X[:, 2], "C": X[:, 3]})
    fit = mg.mcmcglm("Y ~ .", "gaussian", dat, mg.dist_normal(0, 1), w=0.2, n_samples=3000, burnin=300, n_chains=4, seed=11)
    cov = np.linalg.inv(X.T @ X + np.eye(4))
    mu = cov @ X.T @ y
    S = fit.chains[:, 400:, :].reshape(-1, 4)
    se = np.sqrt(np.diag(cov))
    assert np.all(np.abs(S.mean(0) - mu) < 0.1 * se)
    assert np.all(np.abs(S.std(0) / se - 1) < 0.06)
    assert np.all(np.abs(np.quantile(S, 0.975, axis=0) - (mu + 1.96 * se)) < 0.12 * se)


def test_g3_logistic_recovers_truth_and_chains_agree():
    X, y, bt = synth("binomial", 20000, 4, seed=8)
    dat = pd.DataFrame({"Y": y, "A": X[:, 1], "B": X[:, 2], "C": X[:, 3]})
    fit = mg.mcmcglm("Y ~ .", mg.binomial, dat, mg.dist_laplace(0, 1), w=0.1, n_samples=1500, burnin=200, n_chains=4, seed=2)
    S = fit.chains[:, 300:, :]
    m = S.mean(axis=1)                                   # per-chain means
    se = S.reshape(-1, 4).std(0)
    assert np.all(np.abs(m - m.mean(0)) < 0.35 * se)     # chains agree within Monte-Carlo error
    assert np.all(np.abs(m.mean(0) - bt) < 4 * se)       # and sit on the generating coefficients


def test_tuning_sweep_is_one_engine_run_with_per_chain_w():
    """mcmcglm_across_tuningparams (R/slice_utilities.R:43-85): the values of w become the chains of one engine run.
    Each chain must be exactly the chain a single run with that w produces (same Philox substream, same start), and
    the evaluation counts qslice reports per value (dropped by the reference at R/mcmcglm.R:261) come back per value."""
    X, y, _ = synth("binomial", 3000, 4, seed=4)
    m = oracle.make_model("binomial", **PRIOR_CASES["normal"])
    ws = [0.05, 0.2, 0.5, 1.5, 4.0]
    rng = np.random.default_rng(9)
    beta0 = 0.3 * rng.standard_normal((len(ws), 4))
    from mcmcglm_b200 import Engine
    with Engine(3000, 4, family="binomial", w=123.0, n_chains=len(ws), K=8, seed=21, **PRIOR_CASES["normal"]) as e:
        e.set_data(X, y)
        e.set_chain_w(ws)
        for c in range(len(ws)):
            e.init_chain(c, beta0[c])
        S, st = e.run(25)
        per = [e.chain_stats(c) for c in range(len(ws))]
    for c, w in enumerate(ws):
        ref = oracle.run_chain(m, X, y, beta0[c], w=w, n_iter=25, seed=21, chain=c)
        assert np.max(np.abs(S[c] - ref["samples"])) <= 1e-9, w
        assert per[c]["ref_evals"] == ref["n_eval"] and per[c]["stepouts"] == ref["n_stepout"] and per[c]["shrinks"] == ref["n_shrink"]
    assert sum(p_["ref_evals"] for p_ in per) == st["ref_evals"]
    assert per[0]["stepouts"] > per[-1]["stepouts"] and per[-1]["shrinks"] > per[0]["shrinks"]     # narrow w steps out, wide w shrinks
    # and through the front door
    dat = pd.DataFrame({"Y": y, "A": X[:, 1], "B": X[:, 2], "C": X[:, 3]})
    fits = mg.mcmcglm_across_tuningparams(ws, tuning_parameter_name="w", formula="Y ~ .", family="binomial", data=dat,
                                          n_samples=40, burnin=5, seed=3)
    assert [f.w for f in fits] == ws and all(len(mg.samples(f)) == 41 for f in fits)
    assert all(f.stats["nEvaluations"] > 0 for f in fits)
    assert not np.array_equal(mg.samples(fits[0]).iloc[:, :4].to_numpy(), mg.samples(fits[1]).iloc[:, :4].to_numpy())
