"""Jet passes (csrc/cgg_jet.cuh): one pass per update delivers f(x0) and the derivative moments of the
log-likelihood along the coordinate; the decider encloses every candidate's log-potential and only falls back to
exact passes when a comparison is not certain.  Checked here:
  * the enclosure really contains the exact value (40-digit mpmath and the engine's own exact pass), and is tight
    where it matters;
  * chains with and without jet passes are BIT-identical, with identical uniform consumption and identical
    evaluation / step-out / shrink counts (the reference's nEvaluations), for every family x prior, both drivers,
    finite `max`, replayed and Philox uniforms;
  * inflating the error bound (so that some, or all, comparisons fall back to exact passes mid-update) still gives
    the same chain: the hand-over state is exactly the reference algorithm's state at that point.
"""
import numpy as np
import pytest
import mpmath as mp
import oracle
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine

pytestmark = pytest.mark.gpu
mp.mp.dps = 40


def _loglik_mp(family, y, eta, x, delta, sd):
    tot = mp.mpf(0)
    d = mp.mpf(float(delta))
    for yi, ei, xi in zip(y, eta, x):
        t = mp.mpf(float(ei)) + mp.mpf(float(xi)) * d
        if family == "gaussian":
            z = (mp.mpf(float(yi)) - t) / sd
            tot += -z * z / 2 - mp.log(sd) - mp.log(2 * mp.pi) / 2
        elif family == "binomial":
            tot += mp.mpf(float(yi)) * t - mp.log1p(mp.exp(t))
        else:
            tot += mp.mpf(float(yi)) * t - mp.exp(t) - mp.loggamma(mp.mpf(float(yi)) + 1)
    return tot


@pytest.mark.parametrize("family", ["gaussian", "binomial", "poisson"])
def test_enclosure_contains_the_exact_value(family):
    n, p, sd = 2001, 5, 1.3
    X, y, bt = synth(family, n, p, seed=4)
    X[:, 2] *= 7.5                      # a badly scaled column: the per-column power-of-two scaling must cope
    rng = np.random.default_rng(1)
    beta = bt + 0.1 * rng.standard_normal(p)
    beta[2] /= 7.5
    m = oracle.make_model(family, sd=sd, **PRIOR_CASES["normal"])
    with Engine(n, p, family=family, sd=sd, n_chains=1, **PRIOR_CASES["normal"]) as e:
        e.set_data(X, y)
        e.init_chain(0, beta)
        _, eta = e.state(0)
        for j in (0, 2, 4):
            deltas = np.array([0.0, 1e-4, -3e-3, 0.02, -0.05, 0.2, -0.5, 1.0]) / (7.5 if j == 2 else 1.0)
            cands = beta[j] + deltas
            val, bnd, sums = e.debug_jet(0, j, cands)
            f_ex = e.log_potential(0, j, cands)
            for k, dlt in enumerate(cands - beta[j]):
                bb = beta.copy(); bb[j] = cands[k]
                prior = oracle.log_prior_density(m, bb)
                truth = _loglik_mp(family, y, eta, X[:, j], dlt, sd)
                assert np.isfinite(bnd[k]) and bnd[k] >= 0
                assert abs(mp.mpf(float(val[k])) - truth) <= bnd[k], (family, j, k, float(val[k] - truth), bnd[k])
                # and against the engine's own exact pass (what the decider's verdicts must reproduce)
                assert abs(val[k] - (f_ex[k] - prior)) <= bnd[k] + 8 * np.spacing(abs(f_ex[k])), (family, j, k)
            # tight where the slice lives: |x delta| <= ~0.05 => far below any realistic |f - level|
            assert np.all(bnd[:3] < 1e-9), bnd[:3]
            assert np.all(bnd[:5] < 1e-4), bnd[:5]
            if family == "binomial":
                # light pass: no M_0; the enclosure is of the DIFFERENCE f(cand) - f(x0) of two exact evaluations
                dval, dbnd, dsums = e.debug_jet(0, j, cands, light=True)
                assert dsums[0] == 0.0
                base = _loglik_mp(family, y, eta, X[:, j], 0.0, sd)
                for k, dlt in enumerate(cands - beta[j]):
                    truth = _loglik_mp(family, y, eta, X[:, j], dlt, sd) - base
                    assert np.isfinite(dbnd[k]) and abs(mp.mpf(float(dval[k])) - truth) <= dbnd[k], (j, k)
                    # against the difference of the engine's own two exact evaluations (cands[0] is the current point)
                    pk = oracle.log_prior_density(m, np.r_[beta[:j], cands[k], beta[j + 1:]])
                    ex_diff = (f_ex[k] - pk) - (f_ex[0] - oracle.log_prior_density(m, beta))
                    assert abs(dval[k] - ex_diff) <= dbnd[k] + 16 * np.spacing(abs(f_ex[k])), (j, k)
                assert np.all(dbnd[:3] < 1e-7) and np.all(dbnd[:5] < 1e-3)      # order 3 + 2e-11 approximations: looser than a full pass, still << 1


def _run(family, prior, X, y, beta0, iters, U=None, **kw):
    n, p = X.shape
    C = beta0.shape[0]
    with Engine(n, p, family=family, sd=1.0, n_chains=C, **PRIOR_CASES[prior], **kw) as e:
        e.set_data(X, y)
        for c in range(C):
            e.init_chain(c, beta0[c])
        S, st = e.run(iters, replay_u=U)
        states = [e.state(c) for c in range(C)]
    return S, st, states


def _same_chain(a, b):
    (S1, st1, z1), (S2, st2, z2) = a, b
    assert np.array_equal(S1, S2)
    for k in ("uniforms_used", "ref_evals", "stepouts", "shrinks", "updates"):
        assert st1[k] == st2[k], k
    for (b1, e1), (b2, e2) in zip(z1, z2):
        assert np.array_equal(b1, b2) and np.array_equal(e1, e2)


@pytest.mark.parametrize("family,prior,w,max_steps", [
    ("binomial", "laplace", 0.5, -1), ("poisson", "student_t", 0.5, -1), ("gaussian", "normal", 0.05, -1),
    ("binomial", "normal", 0.02, 5), ("poisson", "laplace", 0.01, 3), ("gaussian", "student_t", 0.3, 0),
    ("binomial", "student_t", 3.0, -1), ("gaussian", "laplace", 0.002, -1)])
@pytest.mark.parametrize("driver", ["grid", "cluster", "stepwise"])
def test_jet_chain_is_bit_identical_to_the_exact_chain(family, prior, w, max_steps, driver):
    n, p, C, iters = 3001, 5, 3, 40
    X, y, bt = synth(family, n, p, seed=21)
    rng = np.random.default_rng(8)
    beta0 = 0.5 * rng.standard_normal((C, p))
    U = rng.random((C, 60000))
    for replay in (U, None):
        kw = dict(w=w, max_steps=max_steps, K=6, spec_tau=0.4, driver=driver, seed=77, chain_offset=10)
        exact = _run(family, prior, X, y, beta0, iters, replay, jet=False, **kw)
        assert exact[1]["jet_passes"] == 0
        for light in (True, False):      # binomial: light passes (no M_0) by default; full passes on request
            jet = _run(family, prior, X, y, beta0, iters, replay, jet=True, jet_light=light, **kw)
            _same_chain(exact, jet)
            assert jet[1]["jet_passes"] >= jet[1]["updates"] - jet[1]["jet_fallbacks"] > 0
            if not (light and family == "binomial"):
                assert jet[1]["jet_retries"] == 0
            # inflated bounds: ~1 log unit (some comparisons undecided mid-sequence) and enormous (every one undecided)
            for scale in (1e9, 1e30):
                forced = _run(family, prior, X, y, beta0, iters, replay, jet=True, jet_light=light, jet_bound_scale=scale, **kw)
                _same_chain(exact, forced)
                if scale == 1e30:
                    assert forced[1]["jet_fallbacks"] == forced[1]["updates"]
                    if light and family == "binomial":
                        assert forced[1]["jet_retries"] == forced[1]["updates"]


def test_jet_is_one_pass_per_update_at_scale():
    """n = 1e5 logistic in the stationary regime: essentially every update is decided from its single jet pass."""
    n, p, C = 100_000, 12, 4
    X, y, bt = synth("binomial", n, p, seed=3)
    beta0 = np.tile(bt, (C, 1))
    S, st, _ = _run("binomial", "laplace", X, y, beta0, 25, None, w=0.5, seed=5)
    assert st["updates"] == C * 25 * p
    assert st["jet_fallbacks"] <= 2 and st["jet_retries"] <= 4
    assert st["chain_passes"] <= st["updates"] + 8 * st["jet_fallbacks"] + st["jet_retries"] + C
    exact = _run("binomial", "laplace", X, y, beta0, 25, None, w=0.5, seed=5, jet=False)
    assert np.array_equal(S, exact[0])


def test_rows_at_the_logit_clamp_disable_the_enclosure_not_the_chain():
    # |eta| beyond stats' +-30 clamp: the reference function is not smooth there, the jet pass reports it and the
    # exact passes decide; the chain is the exact chain
    n, p = 4000, 3
    X, y, _ = synth("binomial", n, p, seed=12)
    X[:5, 1] = 40.0
    beta0 = np.array([[0.1, 0.9, -0.2]])
    U = np.random.default_rng(0).random((1, 20000))
    exact = _run("binomial", "normal", X, y, beta0, 15, U, w=0.4, jet=False)
    jet = _run("binomial", "normal", X, y, beta0, 15, U, w=0.4, jet=True)
    _same_chain(exact, jet)
    assert jet[1]["jet_fallbacks"] > 0


def test_light_row_quantities_stay_inside_their_error_budget():
    """The binomial light pass evaluates tanh(|eta|/2), s(1-s) and their product with a 1e-14 exp (64-entry table,
    degree 4) and ONE Newton step on the hardware reciprocal seed; the enclosure budgets JET_LIGHT_EPS = 2e-11 absolute
    for them (cgg_jet.cuh).  Measured over 6.7e7 values of |eta| in [0, 40), both signs."""
    import ctypes as C
    from mcmcglm_b200 import _lib
    err = C.c_double(0.0)
    _lib.check(_lib.load().cgg_debug_light_error(0, C.byref(err)))
    assert 0.0 < err.value <= 1e-11, err.value
