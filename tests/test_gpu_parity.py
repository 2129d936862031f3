"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Gates from the north star:
G1  same (beta, eta, j, candidates) => |f_gpu - f_oracle| <= 1e-12 |f_oracle|
G2  replayed uniforms               => every sample within 1e-9, identical uniform consumption
"""
import json
import os
import numpy as np
import pytest
import oracle
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine, CggError, _lib

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL_G1 = 1e-12
ATOL_G2 = 1e-9


def _engine(family, prior, X, **kw):
    n, p = X.shape
    return Engine(n, p, family=family, sd=kw.pop("sd", 1.0), **PRIOR_CASES[prior], **kw)


@pytest.mark.parametrize("family", ["gaussian", "binomial", "poisson"])
@pytest.mark.parametrize("prior", ["normal", "laplace", "student_t"])
@pytest.mark.parametrize("n", [1001, 100000])
def test_g1_log_potential(family, prior, n):
    p = 7
    X, y, bt = synth(family, n, p, seed=n % 97)
    rng = np.random.default_rng(3)
    beta = bt + 0.2 * rng.standard_normal(p)
    m = oracle.make_model(family, sd=1.3, **PRIOR_CASES[prior])
    with _engine(family, prior, X, sd=1.3) as e:
        e.set_data(X, y)
        e.init_chain(0, beta)
        b_gpu, eta_gpu = e.state(0)
        eta = oracle.init_eta(X, beta)
        assert np.array_equal(b_gpu, beta)
        assert np.array_equal(eta_gpu, eta)          # K4 is bit-identical to the oracle's GEMV
        for j in (0, 3, p - 1):
            cands = beta[j] + np.array([0.0, -0.4, 0.4, 1e-3, -1e-3, 0.05, -0.05, 0.2, 0.11, -0.3, 0.01])
            got = e.log_potential(0, j, cands)       # 11 candidates: two chunks of <= KMAX
            ref = oracle.log_potential(m, X, y, beta, eta, j, cands)
            assert np.all(np.abs(got - ref) <= RTOL_G1 * np.abs(ref)), (got - ref) / ref
        assert abs(e.fx(0) - oracle.log_potential(m, X, y, beta, eta, 0, [beta[0]])[0]) <= RTOL_G1 * abs(e.fx(0))


@pytest.mark.parametrize("sd_eta", [5.0, 15.0, 45.0])
@pytest.mark.parametrize("prior", ["laplace", "normal"])
def test_g1_binomial_large_eta(sd_eta, prior):
    """G1 where the headline workload starts: a chain drawn from the prior at p = 1000 has sd(eta) ~ 45.  For y = 0 and
    large eta R takes log(q) of q = 1 - p with p already rounded (dbinom_raw, R/glm_utils.R:45-47 + stats' logit_linkinv):
    the engine reproduces that form (cgg_math.cuh: rform_log1p_rho), the +-30 clamps included."""
    n, p = 200_000, 6
    rng = np.random.default_rng(int(sd_eta))
    X = np.asfortranarray(rng.standard_normal((n, p)))
    X[:, 0] = 1.0
    beta = rng.standard_normal(p) * sd_eta / np.sqrt(p)
    y = (rng.random(n) < 0.5).astype(np.float64)          # responses unrelated to eta: every (y, sign of eta) pair occurs
    m = oracle.make_model("binomial", **PRIOR_CASES[prior])
    with _engine("binomial", prior, X) as e:
        e.set_data(X, y)
        e.init_chain(0, beta)
        eta = oracle.init_eta(X, beta)
        assert np.std(eta) > 0.8 * sd_eta
        for j in (0, 2, p - 1):
            cands = beta[j] + np.array([0.0, -0.4, 0.4, 1e-3, -2.0, 3.0, 0.05, -0.05])
            got = e.log_potential(0, j, cands)
            ref = oracle.log_potential(m, X, y, beta, eta, j, cands)
            assert np.all(np.abs(got - ref) <= RTOL_G1 * np.abs(ref)), (sd_eta, j, (got - ref) / ref)
        assert abs(e.fx(0) - oracle.log_potential(m, X, y, beta, eta, 0, [beta[0]])[0]) <= RTOL_G1 * abs(e.fx(0))


def test_update_eta_bit_exact():
    X, y, bt = synth("poisson", 5003, 4, seed=5)
    with _engine("poisson", "normal", X) as e:
        e.set_data(X, y)
        e.init_chain(0, bt)
        eta = oracle.init_eta(X, bt)
        for j, nb in [(1, 0.3), (3, -0.2), (1, 0.1)]:
            beta, _ = e.state(0, want_eta=False)
            eta = oracle.update_linear_predictor(nb, beta[j], eta, X[:, j])
            e.update_eta(0, j, nb)
        beta, eta_gpu = e.state(0)
        assert np.array_equal(eta_gpu, eta) and beta[1] == 0.1 and beta[3] == -0.2


@pytest.mark.parametrize("driver", ["grid", "cluster", "stepwise"])
@pytest.mark.parametrize("K,tau", [(1, 0.0), (8, 0.0), (8, 0.5), (3, 0.3)])
def test_g2_readme_golden_replay(driver, K, tau):
    """The reference's README run (seed 42) replayed on the GPU from R's recorded runif stream."""
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    with open(os.path.join(G, "readme_gaussian.json")) as f:
        readme = json.load(f)
    X, y = z["X"], z["y"]
    with Engine(1000, 3, family="gaussian", sd=1.0, prior="normal", prior_mu=0.0, prior_sigma=1.0, w=0.5,
                K=K, spec_tau=tau, driver=driver) as e:
        e.set_data(X, y)
        e.init_chain(0, z["beta0"])
        S, st = e.run(500, replay_u=z["uniforms"])
    S = np.vstack([z["beta0"], S[0]])
    assert np.max(np.abs(S - z["samples"])) <= ATOL_G2
    assert st["uniforms_used"][0] == int(z["uniforms_used"])
    assert st["ref_evals"] == int(z["n_eval"]) and st["stepouts"] == int(z["n_stepout"]) and st["shrinks"] == int(z["n_shrink"])
    assert st["updates"] == 1500
    # and, independently of the oracle, the numbers the reference printed (README.md:79-80, :114-120)
    for row, prow in zip(S[:6], readme["head_samples"]):
        assert np.allclose(row, prow, rtol=2e-7, atol=1e-9)
    coef = S[102:].mean(0)                    # burnin == FALSE rows (quirks Q1/Q3)
    assert np.allclose(coef, list(readme["coef"].values()), rtol=1e-6)


@pytest.mark.parametrize("family,prior,w,max_steps", [
    ("binomial", "laplace", 0.5, -1), ("poisson", "student_t", 0.5, -1), ("gaussian", "normal", 0.05, -1),
    ("binomial", "normal", 0.02, 5), ("poisson", "laplace", 0.01, 3), ("gaussian", "student_t", 0.3, 0)])
@pytest.mark.parametrize("driver", ["grid", "cluster", "stepwise"])
def test_g2_replay_and_philox_vs_oracle(family, prior, w, max_steps, driver):
    n, p, C, iters = 3001, 5, 3, 40
    X, y, bt = synth(family, n, p, seed=21)
    m = oracle.make_model(family, sd=1.0, **PRIOR_CASES[prior])
    rng = np.random.default_rng(8)
    beta0 = 0.5 * rng.standard_normal((C, p))
    U = rng.random((C, 40000))
    for mode in ("replay", "philox"):
        with _engine(family, prior, X, w=w, max_steps=max_steps, n_chains=C, K=6, spec_tau=0.4, driver=driver,
                     seed=77, chain_offset=10) as e:
            e.set_data(X, y)
            for c in range(C):
                e.init_chain(c, beta0[c])
            S, st = e.run(iters, replay_u=U if mode == "replay" else None)
            for c in range(C):
                ref = oracle.run_chain(m, X, y, beta0[c], w=w, n_iter=iters, max_steps=max_steps,
                                       replay_u=U[c] if mode == "replay" else None, seed=77, chain=10 + c)
                assert ref["rc"] == 0
                assert np.max(np.abs(S[c] - ref["samples"])) <= ATOL_G2, (mode, c)
                assert st["uniforms_used"][c] == ref["uniforms_used"]
                beta, eta = e.state(c)
                assert np.max(np.abs(eta - ref["eta"])) <= 1e-9 and np.max(np.abs(beta - ref["beta"])) <= ATOL_G2


@pytest.mark.parametrize("jet", [True, False])
def test_g2_binomial_from_a_laplace_prior_draw(jet):
    """G2 in the transient of the headline workload's shape: beta0 ~ laplace(0, 1) at p = 200 gives sd(eta) ~ 20, i.e.
    thousands of rows beyond stats' +-30 logit clamp and in R's log(1 - p) rounding regime.  Samples, uniform consumption
    and qslice's evaluation / step-out / shrink counts must be the oracle's."""
    n, p, C, iters = 20_000, 200, 2, 3
    X, y, _ = synth("binomial", n, p, seed=31)
    rng = np.random.default_rng(5)
    beta0 = rng.laplace(0.0, 1.0, (C, p))
    m = oracle.make_model("binomial", **PRIOR_CASES["laplace"])
    with _engine("binomial", "laplace", X, w=0.5, n_chains=C, K=8, spec_tau=0.12, seed=11, jet=jet) as e:
        e.set_data(X, y)
        for c in range(C):
            e.init_chain(c, beta0[c])
        assert np.std(e.state(0)[1]) > 12.0
        S, st = e.run(iters)
        ne = ns = nh = 0
        for c in range(C):
            ref = oracle.run_chain(m, X, y, beta0[c], w=0.5, n_iter=iters, seed=11, chain=c)
            assert ref["rc"] == 0
            assert np.max(np.abs(S[c] - ref["samples"])) <= ATOL_G2, c
            assert st["uniforms_used"][c] == ref["uniforms_used"]
            ne += ref["n_eval"]; ns += ref["n_stepout"]; nh += ref["n_shrink"]
            beta, eta = e.state(c)
            assert np.max(np.abs(eta - ref["eta"])) <= 1e-9 * max(1.0, np.max(np.abs(ref["eta"])))
        assert (st["ref_evals"], st["stepouts"], st["shrinks"]) == (ne, ns, nh)


def test_g2_cfg2_shape_full_grid_pair_passes():
    """BASELINE configs[1] at its own size against the oracle: binomial n = 1e5, p = 100, normal prior, 4 chains.  All 147
    worker CTAs take part (slot delivery, 5 slot rounds) and the chains run as pairs; started near the mode so that jet
    passes (light, pair) carry the updates.  2 iterations = 800 updates; the oracle needs ~30 s of CPU."""
    n, p, C, iters = 100_000, 100, 4, 2
    X, y, bt = synth("binomial", n, p, seed=12)
    rng = np.random.default_rng(6)
    beta0 = bt + 0.02 * rng.standard_normal((C, p))
    m = oracle.make_model("binomial", **PRIOR_CASES["normal"])
    with _engine("binomial", "normal", X, w=0.5, n_chains=C, K=8, spec_tau=0.12, seed=3, driver="grid") as e:
        e.set_data(X, y)
        for c in range(C):
            e.init_chain(c, beta0[c])
        ctas, _ = e.launch_shape()
        assert ctas >= 100                      # the full grid, not the few CTAs of the small replay tests
        S, st = e.run(iters)
        assert st["jet_passes"] >= 0.9 * st["updates"]
        ne = 0
        for c in range(C):
            ref = oracle.run_chain(m, X, y, beta0[c], w=0.5, n_iter=iters, seed=3, chain=c)
            assert ref["rc"] == 0
            assert np.max(np.abs(S[c] - ref["samples"])) <= ATOL_G2, c
            assert st["uniforms_used"][c] == ref["uniforms_used"]
            ne += ref["n_eval"]
        assert st["ref_evals"] == ne


def test_g2_cfg2_shape_cluster_driver():
    """The same workload on the driver it gets by default (n <= 2^18: one thread-block cluster per chain, sums exchanged
    through distributed shared memory, every CTA of a cluster decides for itself)."""
    n, p, C, iters = 100_000, 100, 4, 2
    X, y, bt = synth("binomial", n, p, seed=12)
    rng = np.random.default_rng(6)
    beta0 = bt + 0.02 * rng.standard_normal((C, p))
    m = oracle.make_model("binomial", **PRIOR_CASES["normal"])
    with _engine("binomial", "normal", X, w=0.5, n_chains=C, K=8, spec_tau=0.12, seed=3) as e:
        e.set_data(X, y)
        for c in range(C):
            e.init_chain(c, beta0[c])
        ctas, _ = e.launch_shape()
        assert ctas % C == 0 and 2 <= ctas // C <= 16        # C clusters
        S, st = e.run(iters)
        assert st["jet_passes"] >= 0.9 * st["updates"] and st["launches"] == 1
        for c in (0, 3):
            ref = oracle.run_chain(m, X, y, beta0[c], w=0.5, n_iter=iters, seed=3, chain=c)
            assert ref["rc"] == 0
            assert np.max(np.abs(S[c] - ref["samples"])) <= ATOL_G2, c
            assert st["uniforms_used"][c] == ref["uniforms_used"]
            assert e.chain_stats(c)["ref_evals"] == ref["n_eval"]
            beta, eta = e.state(c)
            assert np.max(np.abs(eta - ref["eta"])) <= 1e-9


def test_chunked_runs_continue_the_chain():
    X, y, bt = synth("binomial", 2000, 4, seed=2)
    m = oracle.make_model("binomial", **PRIOR_CASES["normal"])
    ref = oracle.run_chain(m, X, y, np.zeros(4), w=0.3, n_iter=30, seed=5, chain=0)
    with _engine("binomial", "normal", X, w=0.3, seed=5) as e:
        e.set_data(X, y)
        e.init_chain(0, np.zeros(4))
        parts = [e.run(k)[0][0] for k in (7, 13, 10)]
    assert np.max(np.abs(np.vstack(parts) - ref["samples"])) <= ATOL_G2


def test_large_n_roundtrip_properties():
    """Size-independent checks at a BASELINE-sized column (n = 1e6): the committed eta always equals
    X beta recomputed from scratch, and f(x0) carried by the sweep equals a fresh evaluation."""
    n, p = 1_000_000, 6
    X, y, bt = synth("binomial", n, p, seed=9)
    with _engine("binomial", "laplace", X, w=0.5, n_chains=2, K=8) as e:
        e.set_data(X, y)
        e.init_chain(0, np.zeros(p))
        e.init_chain(1, 0.1 * np.ones(p))
        S, st = e.run(3)
        for c in range(2):
            beta, eta = e.state(c)
            assert np.array_equal(beta, S[c, -1])
            assert np.max(np.abs(eta - X @ beta)) < 1e-11
            carried = e.fx(c)
            fresh = e.log_potential(c, 0, [beta[0]])[0]
            assert abs(carried - fresh) <= 1e-12 * abs(fresh)
    assert st["updates"] == 2 * 3 * p


def test_errors():
    X, y, _ = synth("binomial", 100, 2, seed=1)
    with _engine("binomial", "normal", X) as e:
        with pytest.raises(CggError) as ei:
            e.run(1)
        assert ei.value.code == _lib.E_STATE
        ybad = y.copy(); ybad[3] = 2.0
        with pytest.raises(CggError) as ei:
            e.set_data(X, ybad)
        assert ei.value.code == _lib.E_ARG
        e.set_data(X, y)
        with pytest.raises(CggError) as ei:
            e.run(1)
        assert ei.value.code == _lib.E_STATE      # chain not initialised
        e.init_chain(0, np.zeros(2))
        with pytest.raises(CggError) as ei:
            e.run(5, replay_u=np.full(7, 0.5))
        assert ei.value.code == _lib.E_STREAM
        e.init_chain(0, np.array([np.nan, 0.0]))
        with pytest.raises(CggError) as ei:
            e.run(1)
        assert ei.value.code == _lib.E_NAN


def test_neg_inf_is_outside_the_slice_not_an_error():
    # poisson: huge candidate => exp overflow => log-potential -Inf => "y < -Inf" is FALSE (R semantics)
    X, y, _ = synth("poisson", 500, 2, seed=1)
    m = oracle.make_model("poisson", **PRIOR_CASES["normal"])
    with _engine("poisson", "normal", X, w=2000.0, seed=3) as e:
        e.set_data(X, y)
        e.init_chain(0, np.zeros(2))
        f = e.log_potential(0, 1, [1500.0, -1500.0, 0.01])
        fr = oracle.log_potential(m, X, y, np.zeros(2), np.zeros(500), 1, [1500.0, -1500.0, 0.01])
        assert f[0] == -np.inf and fr[0] == -np.inf and f[1] == fr[1] and np.isfinite(f[2])
        S, _ = e.run(5)
    ref = oracle.run_chain(m, X, y, np.zeros(2), w=2000.0, n_iter=5, seed=3, chain=0)
    assert np.max(np.abs(S[0] - ref["samples"])) <= ATOL_G2
