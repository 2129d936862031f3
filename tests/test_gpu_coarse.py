"""The fp32 pre-filter must never change a result: it may only (a) reject a candidate that is outside the
slice for certain, or (b) leave it undecided and have it re-scored in fp64."""
import ctypes as C
import os
import numpy as np
import pytest
import oracle
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine, _lib

pytestmark = pytest.mark.gpu


def test_fp32_softplus_error_constant_holds_for_every_float():
    L = _lib.load()
    a, b = C.c_double(), C.c_double()
    _lib.check(L.cgg_debug_coarse_error(0, C.byref(a), C.byref(b)))
    kappa = 2.0 ** -21                       # kCoarseKappa in cgg_math.cuh
    assert 0 < a.value <= kappa / 2, (a.value, b.value)     # >= 2x margin over the exhaustive maximum


def _run(X, y, beta0, prefilter, w=0.5, iters=30, seed=4, theta=None, **kw):
    if theta is not None:
        os.environ["CGG_COARSE_THETA"] = str(theta)
    else:
        os.environ.pop("CGG_COARSE_THETA", None)
    try:
        n, p = X.shape
        kw.setdefault("driver", "grid")     # the pre-filter belongs to the grid-wide kernels (the cluster driver scores everything in fp64)
        with Engine(n, p, family="binomial", w=w, n_chains=beta0.shape[0], K=8, seed=seed, prefilter=prefilter, jet=False,   # this file tests the exact-pass path
                    **PRIOR_CASES["laplace"], **kw) as e:
            e.set_data(X, y)
            for c in range(beta0.shape[0]):
                e.init_chain(c, beta0[c])
            S, st = e.run(iters)
        return S, st
    finally:
        os.environ.pop("CGG_COARSE_THETA", None)


@pytest.mark.parametrize("n", [20001, 300000])
def test_prefilter_is_bit_identical_to_all_fp64(n):
    X, y, bt = synth("binomial", n, 6, seed=31)
    beta0 = np.random.default_rng(1).standard_normal((3, 6))
    S0, st0 = _run(X, y, beta0, prefilter=False)
    S1, st1 = _run(X, y, beta0, prefilter=True)
    assert st0["coarse_evals"] == 0 and st1["coarse_evals"] > 0.3 * st1["cand_evals"]
    assert np.array_equal(S0, S1)
    for k in ("ref_evals", "stepouts", "shrinks", "updates", "uniforms_used"):
        assert st0[k] == st1[k]
    # an absurdly aggressive policy (everything further than 0.01 slice widths is pre-filtered) leaves many
    # candidates undecided -- and still changes nothing
    S2, st2 = _run(X, y, beta0, prefilter=True, theta=0.01)
    assert st2["coarse_undecided"] > 0 and np.array_equal(S0, S2)
    # and the chain is the oracle's
    m = oracle.make_model("binomial", **PRIOR_CASES["laplace"])
    if n < 50000:
        ref = oracle.run_chain(m, X, y, beta0[1], w=0.5, n_iter=30, seed=4, chain=1)
        assert np.max(np.abs(S1[1] - ref["samples"])) <= 1e-9


def test_prefilter_with_rows_on_the_logit_clamp():
    # huge coefficients put many rows beyond |eta| = 30 and some within 0.02 of it for some candidate
    rng = np.random.default_rng(5)
    n, p = 40000, 3
    X = np.asfortranarray(rng.standard_normal((n, p)) * 12.0)
    X[:, 0] = 1.0
    X[:200, 1] = 30.0 / 2.5                     # eta = x * beta exactly at the threshold when beta_1 = 2.5
    y = (rng.random(n) < 0.5).astype(float)
    beta0 = np.array([[0.0, 2.5, -1.0], [0.3, 2.4999, 1.0]])
    S0, _ = _run(X, y, beta0, prefilter=False, w=0.05, iters=15)
    S1, st1 = _run(X, y, beta0, prefilter=True, w=0.05, iters=15, theta=0.01)
    assert np.array_equal(S0, S1) and st1["coarse_evals"] > 0
