"""CPU oracle for the mcmcglm CGGibbs hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  mcmcglm_b200 (the product) must never import it.

`oracle.c` is the restatement (one function per reference/external behaviour, each citing
file:line); this module is a thin ctypes binding plus `r_rng.py` (R's Mersenne-Twister, used to
regenerate the README seed-42 example that pins the oracle to the reference's printed output).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

GAUSSIAN, BINOMIAL, POISSON, NEGBIN, BINOMIAL_PROBIT = 0, 1, 2, 3, 4
NORMAL, LAPLACE, STUDENT_T, GAMMA, EXPONENTIAL = 0, 1, 2, 3, 4
FAMILIES = {"gaussian": GAUSSIAN, "binomial": BINOMIAL, "poisson": POISSON, "negative_binomial": NEGBIN, "binomial_probit": BINOMIAL_PROBIT}
PRIORS = {"normal": NORMAL, "laplace": LAPLACE, "student_t": STUDENT_T, "gamma": GAMMA, "exponential": EXPONENTIAL}
MAX_PRIORS = 8
OK, E_NAN, E_STREAM, E_NOTERM, E_ARG = 0, -1, -2, -3, -4


class Model(C.Structure):
    _fields_ = [("family", C.c_int), ("sd", C.c_double), ("prior", C.c_int),
                ("pmu", C.c_double), ("psigma", C.c_double), ("pdf", C.c_double),
                ("n_more", C.c_int), ("more_prior", C.c_int * (MAX_PRIORS - 1)),
                ("more_a", C.c_double * (MAX_PRIORS - 1)), ("more_b", C.c_double * (MAX_PRIORS - 1)), ("more_c", C.c_double * (MAX_PRIORS - 1))]


class SliceStats(C.Structure):
    _fields_ = [("n_eval", C.c_int64), ("n_stepout", C.c_int64), ("n_shrink", C.c_int64)]


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        for name in ("orc_stirlerr",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.orc_pnorm.restype = C.c_double
        L.orc_pnorm.argtypes = [C.c_double]
        L.orc_dnbinom_mu_log.restype = C.c_double
        L.orc_dnbinom_mu_log.argtypes = [C.c_double] * 3
        L.orc_dgamma_log.restype = C.c_double
        L.orc_dgamma_log.argtypes = [C.c_double] * 3
        for name in ("orc_bd0", "orc_dpois_log", "orc_dt_log", "orc_dexp_log"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double, C.c_double]
        for name in ("orc_dnorm_log", "orc_dbinom_log"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double] * 3
        L.orc_log_density.restype = C.c_double
        L.orc_log_density.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_prior_log_density1.restype = C.c_double
        L.orc_prior_log_density1.argtypes = [C.POINTER(Model), C.c_double]
        L.orc_log_prior_density.restype = C.c_double
        L.orc_log_prior_density.argtypes = [C.POINTER(Model), C.c_int64, _dp]
        L.orc_linkinv.restype = None
        L.orc_linkinv.argtypes = [C.c_int, C.c_int64, _dp, _dp]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_init_eta.restype = None
        L.orc_init_eta.argtypes = [C.c_int64, C.c_int64, _dp, C.c_int64, _dp, _dp]
        L.orc_update_linear_predictor.restype = None
        L.orc_update_linear_predictor.argtypes = [C.c_int64, C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_log_potential_batch.restype = C.c_int
        L.orc_log_potential_batch.argtypes = [C.POINTER(Model), C.c_int64, C.c_int64, _dp, C.c_int64, _dp,
                                              _dp, _dp, C.c_int64, C.c_int, _dp, _dp]
        L.orc_log_potential_naive.restype = C.c_double
        L.orc_log_potential_naive.argtypes = [C.POINTER(Model), C.c_int64, C.c_int64, _dp, C.c_int64, _dp,
                                              _dp, C.c_int64, C.c_double, _dp]
        L.orc_run_chain.restype = C.c_int
        L.orc_run_chain.argtypes = [C.POINTER(Model), C.c_int64, C.c_int64, _dp, C.c_int64, _dp, _dp, _dp,
                                    C.c_double, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_uint64,
                                    C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, _dp,
                                    C.POINTER(C.c_uint64), C.POINTER(SliceStats)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_model(family="gaussian", sd=1.0, prior="normal", prior_mu=0.0, prior_sigma=1.0, prior_df=1.0, more_priors=()):
    """more_priors: further components of a LIST of priors, each (kind, a, b, c) with (a, b, c) = (mu, sigma, df) for
    normal / laplace / student_t, (shape, rate, -) for gamma, (-, rate, -) for exponential."""
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    pri = PRIORS[prior] if isinstance(prior, str) else int(prior)
    m = Model(fam, float(sd), pri, float(prior_mu), float(prior_sigma), float(prior_df))
    m.n_more = len(more_priors)
    for k, (kind, a, b, c) in enumerate(more_priors):
        m.more_prior[k] = PRIORS[kind] if isinstance(kind, str) else int(kind)
        m.more_a[k], m.more_b[k], m.more_c[k] = float(a), float(b), float(c)
    return m


def _colmajor(X):
    """Return (buffer, ld) for an n x p matrix as R stores it (column-major, ld = n)."""
    X = np.asarray(X, dtype=np.float64)
    return np.asfortranarray(X), X.shape[0]


def linkinv(family, eta):
    eta = _f64(eta)
    mu = np.empty_like(eta)
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    lib().orc_linkinv(fam, eta.size, _p(eta), _p(mu))
    return mu


def log_density(family, mu, y, sd=1.0):
    fam = FAMILIES[family] if isinstance(family, str) else int(family)
    mu = np.broadcast_to(_f64(mu), np.shape(y)).ravel()
    y = _f64(y).ravel()
    return np.array([lib().orc_log_density(fam, float(m), float(v), float(sd)) for m, v in zip(mu, y)])


def log_prior_density(model, beta):
    beta = _f64(beta)
    return lib().orc_log_prior_density(C.byref(model), beta.size, _p(beta))


def init_eta(X, beta):
    Xf, ld = _colmajor(X)
    n, p = Xf.shape
    beta = _f64(beta)
    eta = np.empty(n)
    lib().orc_init_eta(n, p, _p(Xf), ld, _p(beta), _p(eta))
    return eta


def update_linear_predictor(new_beta_j, current_beta_j, current_eta, X_j):
    """R/glm_utils.R:126-132"""
    eta = _f64(current_eta)
    xj = _f64(X_j)
    out = np.empty_like(eta)
    lib().orc_update_linear_predictor(eta.size, float(new_beta_j), float(current_beta_j), _p(eta), _p(xj), _p(out))
    return out


def log_potential(model, X, y, beta, eta, j, cands):
    """R/glm_utils.R:187-218 ("update") at each candidate new_beta_j; j is 0-based."""
    Xf, ld = _colmajor(X)
    n, p = Xf.shape
    y, beta, eta = _f64(y), _f64(beta), _f64(eta)
    cands = np.atleast_1d(_f64(cands))
    out = np.empty(cands.size)
    rc = lib().orc_log_potential_batch(C.byref(model), n, p, _p(Xf), ld, _p(y), _p(beta), _p(eta), int(j),
                                       cands.size, _p(cands), _p(out))
    assert rc == 0
    return out


def log_potential_naive(model, X, y, beta, j, cand):
    Xf, ld = _colmajor(X)
    n, p = Xf.shape
    y, beta = _f64(y), _f64(beta)
    scratch = np.empty(2 * n + p)
    return lib().orc_log_potential_naive(C.byref(model), n, p, _p(Xf), ld, _p(y), _p(beta), int(j), float(cand),
                                         _p(scratch))


def philox_uniform(seed, chain, idx):
    return lib().orc_philox_uniform(int(seed), int(chain), int(idx))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def run_chain(model, X, y, beta0, w, n_iter, max_steps=-1, replay_u=None, seed=0, chain=0, stream_pos0=0,
              max_updates=0, compute_mu=False, eta0=None):
    """R/mcmcglm.R:226-274.  Returns dict(samples [n_iter x p], beta, eta, uniforms_used, stats, rc)."""
    Xf, ld = _colmajor(X)
    n, p = Xf.shape
    y = _f64(y)
    beta = _f64(beta0).copy()
    eta = init_eta(Xf, beta) if eta0 is None else _f64(eta0).copy()
    samples = np.full((n_iter, p), np.nan)
    used = C.c_uint64(0)
    stats = SliceStats()
    if replay_u is not None:
        ru = _f64(replay_u)
        ru_p, n_u = _p(ru), ru.size
    else:
        ru_p, n_u = None, 0
    rc = lib().orc_run_chain(C.byref(model), n, p, _p(Xf), ld, _p(y), _p(beta), _p(eta), float(w), int(max_steps),
                             int(n_iter), int(max_updates), ru_p, n_u, int(seed), int(chain), int(stream_pos0),
                             int(bool(compute_mu)), _p(samples), C.byref(used), C.byref(stats))
    return dict(samples=samples, beta=beta, eta=eta, uniforms_used=used.value, rc=rc,
                n_eval=stats.n_eval, n_stepout=stats.n_stepout, n_shrink=stats.n_shrink)
