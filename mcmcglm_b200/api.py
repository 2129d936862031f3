"""Host-side mirror of the reference's public interface for the CGGibbs path.

The drop-in boundary of this project is the C ABI (include/cggibbs.h); the host language of the reference
is R, whose glue lives under r/ (it cannot be executed in this image: R is not installed).  This module
restates the same interface -- same names, argument meaning, defaults, error behaviour and object layout --
in Python over the same C ABI, so that the parity tests read like tests of the reference:

    mcmcglm(formula, family, data, beta_prior, log_likelihood_extra_args, linear_predictor_calc,
            sample_method, qslice_fun, ..., n_samples, burnin)          R/mcmcglm.R:147-157
    samples(x) / coef(x) / quantile(x, probs) / print                    R/mcmcglm_methods.R:2-158
    log_potential_from_betaj(...)                                        R/glm_utils.R:187-218
    update_linear_predictor(...)                                         R/glm_utils.R:126-132
    mcmcglm_across_tuningparams(...)                                     R/slice_utilities.R:43-85

Everything numeric runs on the GPU through libcggibbs.so.  What the engine does not implement is rejected
with the reason (no CPU fallback): other families/links, non-iid or other priors, other qslice samplers,
sample_method = "normal-normal", linear_predictor_calc = "naive".
"""
import re
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .engine import Engine


# ------------------------------------------------------------------------------ priors (distributional::)
@dataclass(frozen=True)
class Dist:
    kind: str
    mu: float = 0.0
    sigma: float = 1.0
    df: float = 1.0

    def generate(self, n, rng):
        """distributional::generate(dist, n): the draw mcmcglm() uses for init_beta (R/mcmcglm.R:208)."""
        if self.kind == "normal":
            return self.mu + self.sigma * rng.standard_normal(n)
        if self.kind == "laplace":
            return rng.laplace(self.mu, self.sigma, n)
        return self.mu + self.sigma * rng.standard_t(self.df, n)


def dist_normal(mu=0.0, sigma=1.0):
    return Dist("normal", float(mu), float(sigma))


def dist_laplace(mu=0.0, sigma=1.0):
    return Dist("laplace", float(mu), float(sigma))


def dist_student_t(df, mu=0.0, sigma=1.0):
    return Dist("student_t", float(mu), float(sigma), float(df))


# ------------------------------------------------------------------------------ families (stats::)
@dataclass(frozen=True)
class Family:
    family: str
    link: str


def gaussian(link="identity"):
    return Family("gaussian", link)


def binomial(link="logit"):
    return Family("binomial", link)


def poisson(link="log"):
    return Family("poisson", link)


_FAMILY_BY_NAME = {"gaussian": gaussian, "binomial": binomial, "poisson": poisson}


def check_family(family):
    """R/family_data_processing.R:3-16: a name, a family function or a family object."""
    if isinstance(family, str):
        if family not in _FAMILY_BY_NAME:
            raise L.CggError(L.E_UNSUPPORTED, f"family {family!r} is not supported by the GPU engine "
                             f"(supported: {sorted(_FAMILY_BY_NAME)})")
        family = _FAMILY_BY_NAME[family]
    if callable(family):
        family = family()
    if getattr(family, "family", None) is None:
        raise ValueError("'family' not recognized")
    return family


# ------------------------------------------------------------------------------ the slice sampler plug-in
def slice_stepping_out(*a, **k):
    """Stand-in for qslice::slice_stepping_out.  A GPU kernel cannot call back into a host closure, so the
    plug-in protocol of R/mcmcglm.R:258-261 is honoured by identity: passing THIS function as qslice_fun
    selects the engine's built-in stepping-out/shrinkage sampler (tuning arguments `w`, `max`).  It is not
    callable on the host."""
    raise RuntimeError("slice_stepping_out is executed on the GPU by the engine; it cannot be called directly")


# ------------------------------------------------------------------------------ model frame
def extract_model_data(formula, data):
    """R/family_data_processing.R:20-36 for the formulas the tests use: `Y ~ .`, `Y ~ a + b`, optional
    `- 1` / `+ 0`.  Returns (Y, X, column names); X has the `(Intercept)` column first like model.matrix."""
    import pandas as pd
    if not isinstance(data, pd.DataFrame):
        data = pd.DataFrame(data)
    m = re.fullmatch(r"\s*([A-Za-z_.][\w.]*)\s*~\s*(.+?)\s*", formula)
    if not m:
        raise ValueError(f"cannot parse formula {formula!r}")
    resp, rhs = m.group(1), m.group(2)
    if resp not in data.columns:
        raise KeyError(f"response {resp!r} not found in data")
    intercept = True
    terms = []
    for sign, tok in re.findall(r"([+-]?)\s*([^+-]+)", rhs):
        tok = tok.strip()
        if tok in ("0", "1"):
            if (tok == "0") or (sign == "-" and tok == "1"):
                intercept = False
            continue
        if sign == "-":
            raise ValueError(f"unsupported formula term '- {tok}'")
        if tok == ".":
            terms += [c for c in data.columns if c != resp and c not in terms]
        elif re.fullmatch(r"[A-Za-z_.][\w.]*", tok):
            if tok not in data.columns:
                raise KeyError(f"variable {tok!r} not found in data")
            terms.append(tok)
        else:
            raise ValueError(f"unsupported formula term {tok!r} (only main effects of numeric columns)")
    cols, names = [], []
    if intercept:
        cols.append(np.ones(len(data)))
        names.append("(Intercept)")
    for t in terms:
        cols.append(np.asarray(data[t], dtype=np.float64))
        names.append(t)
    X = np.asfortranarray(np.column_stack(cols)) if cols else np.zeros((len(data), 0), order="F")
    return np.asarray(data[resp], dtype=np.float64), X, names


# ------------------------------------------------------------------------------ the returned object
@dataclass
class McmcGlm:
    """The list of R/mcmcglm.R:282-297 with class c("mcmcglm", "list")."""
    beta_samples: "object"      # DataFrame: one column per coefficient + iteration + burnin (quirk Q1)
    beta_mean: "object"         # 1-row DataFrame over burnin == False rows (quirk Q3)
    data: "object"
    model_matrix: np.ndarray
    param_list: "object"        # None: the per-iteration (beta, eta, mu) store is not materialised
    family: Family
    formula: str
    call: str
    burnin: int
    sample_method: str
    qslice_fun: "object"
    tuning: dict = field(default_factory=dict)      # the `...` of the call, e.g. w (quirk Q12)
    chains: np.ndarray = None                       # extension: [n_chains, n_samples + 1, p]
    stats: dict = None                              # extension: engine counters (qslice's nEvaluations etc.)

    def __getattr__(self, name):                    # x$w like the reference's `c(list(...), list(...))`
        t = self.__dict__.get("tuning", {})
        if name in t:
            return t[name]
        raise AttributeError(name)

    def __repr__(self):                             # print.mcmcglm, R/mcmcglm_methods.R:2-9
        return ("Object of class 'mcmcglm'\n\nCall:  " + self.call + "\n\nAverage of parameter samples:\n"
                + self.beta_mean.to_string() + "\n")


def samples(x):
    """samples.mcmcglm, R/mcmcglm_methods.R:48-50"""
    return x.beta_samples


def coef(x):
    """coef.mcmcglm, R/mcmcglm_methods.R:84-86"""
    return x.beta_mean


def quantile(x, probs=(0.025, 0.5, 0.975)):
    """quantile.mcmcglm, R/mcmcglm_methods.R:124-158.  Like the reference it summarises the rows flagged
    burnin == TRUE (quirk Q2) and adds a `mean` column; quantiles are R's default type 7."""
    import pandas as pd
    n_vars = x.model_matrix.shape[1]
    col_names = ["q_" + _r_num(p).replace(".", "") for p in probs]    # paste("q_", gsub("\\.", "", probs))
    S = samples(x)
    B = S[S["burnin"]].iloc[:, :n_vars]
    rows = []
    for var in B.columns:
        v = B[var].to_numpy()
        rows.append([var, v.mean()] + [np.quantile(v, p) for p in probs])
    return pd.DataFrame(rows, columns=["var", "mean"] + col_names)


def _r_num(p):
    s = repr(float(p))
    return s[:-2] if s.endswith(".0") else s


# ------------------------------------------------------------------------------ the front door
def _engine_kwargs(family, beta_prior, log_likelihood_extra_args):
    fam = check_family(family)
    if fam.family not in ("gaussian", "binomial", "poisson"):
        raise L.CggError(L.E_UNSUPPORTED, f"family {fam.family!r} is not supported by the GPU engine")
    if isinstance(beta_prior, (list, tuple)):
        raise L.CggError(L.E_UNSUPPORTED, "a list of per-coordinate priors is not supported by the GPU engine "
                         "(only one iid prior: dist_normal, dist_laplace, dist_student_t)")
    if not isinstance(beta_prior, Dist):
        raise L.CggError(L.E_UNSUPPORTED, f"prior {beta_prior!r} is not supported by the GPU engine "
                         "(supported: dist_normal, dist_laplace, dist_student_t)")
    sd = float((log_likelihood_extra_args or {}).get("sd", 1.0))
    return fam, dict(family=fam.family, link=fam.link, sd=sd, prior=beta_prior.kind, prior_mu=beta_prior.mu,
                     prior_sigma=beta_prior.sigma, prior_df=beta_prior.df)


def mcmcglm(formula, family="gaussian", data=None, beta_prior=None, log_likelihood_extra_args=None,
            linear_predictor_calc="update", sample_method="slice_sampling", qslice_fun=slice_stepping_out,
            n_samples=500, burnin=100, *, n_chains=1, device=0, K=8, seed=None, beta_init=None,
            replay_uniforms=None, driver="persistent", _w_per_chain=None, **tuning):
    """mcmcglm() of R/mcmcglm.R:147-299 on the GPU engine.  `**tuning` is the reference's `...` (forwarded
    to qslice_fun: `w`, `max`).  Keyword-only arguments after `burnin` are engine extensions.
    """
    if beta_prior is None:
        beta_prior = dist_normal(0, 1)                                        # :150
    if log_likelihood_extra_args is None:
        log_likelihood_extra_args = {"sd": 1}                                 # :151
    if linear_predictor_calc not in ("update", "naive"):                      # match.arg, :161
        raise ValueError("'arg' should be one of 'update', 'naive'")
    if sample_method not in ("slice_sampling", "normal-normal"):              # match.arg, :163
        raise ValueError("'arg' should be one of 'slice_sampling', 'normal-normal'")
    if burnin >= n_samples:                                                   # :165
        raise ValueError("Need more iterations than burnin")
    if len(tuning) == 0 and sample_method == "slice_sampling":                # :167-169
        raise ValueError("A tuning parameter for the `qslice_fun` is missing. For default choice of "
                         "`qslice::slice_stepping_out` a slice width w needs to be provided")
    if sample_method == "normal-normal":
        raise L.CggError(L.E_UNSUPPORTED, "sample_method = 'normal-normal' (the reference's closed-form test "
                         "sampler, R/sampling.R) is not part of the GPU path")
    if linear_predictor_calc == "naive":
        raise L.CggError(L.E_UNSUPPORTED, "linear_predictor_calc = 'naive' is not part of the GPU path "
                         "(the engine always uses the O(n) CGGibbs update)")
    if qslice_fun is not slice_stepping_out:
        raise L.CggError(L.E_UNSUPPORTED, "only qslice::slice_stepping_out is implemented on the GPU; "
                         f"got {getattr(qslice_fun, '__name__', qslice_fun)!r}")
    unknown = set(tuning) - {"w", "max"}
    if unknown or "w" not in tuning:
        raise L.CggError(L.E_ARG, f"slice_stepping_out takes the tuning arguments w (required) and max; got {sorted(tuning)}")
    fam, ekw = _engine_kwargs(family, beta_prior, log_likelihood_extra_args)
    Y, X, names = extract_model_data(formula, data)                           # :176-178
    n, p = X.shape
    rng = np.random.default_rng(seed)
    if beta_init is None:                                                     # :200-213
        beta0 = np.stack([beta_prior.generate(p, rng) for _ in range(n_chains)])
    else:
        beta0 = np.broadcast_to(np.asarray(beta_init, dtype=np.float64), (n_chains, p)).copy()
    mx = tuning.get("max", np.inf)
    eng_seed = int(rng.integers(0, 2 ** 63 - 1)) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 63))
    with Engine(n, p, w=float(tuning["w"]), max_steps=-1 if np.isinf(mx) else int(mx), n_chains=n_chains, K=K,
                device=device, driver=driver, seed=eng_seed, **ekw) as e:
        e.set_data(X, Y)
        if _w_per_chain is not None:                                          # a tuning sweep: every chain its own w
            e.set_chain_w(_w_per_chain)
        for c in range(n_chains):
            e.init_chain(c, beta0[c])                                         # :215 init_eta = X %*% init_beta
        S, st = e.run(n_samples, replay_u=replay_uniforms)                    # :226-274
        st["per_chain"] = [e.chain_stats(c) for c in range(n_chains)]
    import pandas as pd
    chains = np.concatenate([beta0[:, None, :], S], axis=1)                   # row 0 = the prior draw, :222
    df = pd.DataFrame(chains[0], columns=names)
    df["iteration"] = np.arange(n_samples + 1)
    df["burnin"] = df["iteration"] <= burnin + 1                              # :197-198 (quirk Q1)
    beta_mean = df.loc[~df["burnin"], names].mean().to_frame().T              # :276-280 (quirk Q3)
    call = (f"mcmcglm(formula = {formula}, family = \"{fam.family}\", data = <data>, beta_prior = {beta_prior.kind}"
            f"({beta_prior.mu:g}, {beta_prior.sigma:g}), " + ", ".join(f"{k} = {v}" for k, v in tuning.items()) + ")")
    return McmcGlm(beta_samples=df, beta_mean=beta_mean, data=data, model_matrix=X, param_list=None, family=fam,
                   formula=formula, call=call, burnin=burnin, sample_method=sample_method, qslice_fun=qslice_fun,
                   tuning=dict(tuning), chains=chains, stats=st)


# ------------------------------------------------------------------------------ operators
def log_potential_from_betaj(new_beta_j, j, current_beta, current_eta, Y, X, family, beta_prior,
                             linear_predictor_calc="update", device=0, **extra):
    """log_potential_from_betaj of R/glm_utils.R:187-218 evaluated by the GPU kernel (K1).
    `j` is 1-based, as in the reference.  `new_beta_j` may be a scalar or an array of candidates."""
    if linear_predictor_calc != "update":
        raise L.CggError(L.E_UNSUPPORTED, "linear_predictor_calc = 'naive' is not part of the GPU path")
    _, ekw = _engine_kwargs(family, beta_prior, {"sd": extra.get("sd", 1.0)})
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    if not 1 <= j <= p:
        raise IndexError("j is 1-based and must be in 1..ncol(X)")
    with Engine(n, p, w=1.0, n_chains=1, device=device, driver="stepwise", **ekw) as e:
        e.set_data(X, Y)
        e.set_state(0, current_beta, current_eta)
        out = e.log_potential(0, j - 1, new_beta_j)
    return float(out[0]) if np.ndim(new_beta_j) == 0 else out


def update_linear_predictor(new_beta_j, current_beta_j, current_eta, X_j, device=0):
    """update_linear_predictor of R/glm_utils.R:126-132 evaluated by the GPU kernel (K2): two roundings."""
    xj = np.asarray(X_j, dtype=np.float64).reshape(-1, 1)
    n = xj.shape[0]
    with Engine(n, 1, family="gaussian", w=1.0, n_chains=1, device=device, driver="stepwise") as e:
        e.set_data(xj, np.zeros(n))
        e.set_state(0, [current_beta_j], current_eta)
        e.update_eta(0, 0, new_beta_j)
        return e.state(0)[1]


def mcmcglm_across_tuningparams(*values, tuning_parameter_name="w", parallelise=False, n_cores=None, **kw):
    """mcmcglm_across_tuningparams of R/slice_utilities.R:43-85.  The reference runs one mcmcglm() per value of the tuning
    parameter (`lapply`, or `future_lapply` over worker processes with `parallelise = TRUE`).  Here a sweep over `w` is
    ONE engine run: the values become the chains of a single upload of X and y (every chain its own `w`, prior draw and
    Philox substream), walking the columns together, up to 32 values per run.  `parallelise` / `n_cores` are accepted and
    ignored.  Returns the list of mcmcglm objects (one per value, in order), each with `stats["nEvaluations"]` = the number
    of log-potential evaluations qslice made for that value -- which the reference computes and drops (R/mcmcglm.R:261)."""
    if len(values) == 0:
        raise ValueError("a vector of tuning parameter values is needed")
    vals = list(np.atleast_1d(values[0])) if np.ndim(values[0]) else list(values)
    other = {}
    if np.ndim(values[0]) and len(values) > 1:                      # R: further positional tuning args are passed through
        raise ValueError("pass further tuning parameters by name")
    if tuning_parameter_name != "w":
        return [mcmcglm(**{tuning_parameter_name: v}, **kw) for v in vals]
    kw = dict(kw)
    kw.pop("w", None)
    user_chains = kw.pop("n_chains", 1)
    if user_chains != 1:
        raise L.CggError(L.E_ARG, "mcmcglm_across_tuningparams runs one chain per tuning value")
    out = []
    for i0 in range(0, len(vals), 32):
        chunk = [float(v) for v in vals[i0:i0 + 32]]
        fit = mcmcglm(w=chunk[0], n_chains=len(chunk), _w_per_chain=np.array(chunk), **kw, **other)
        for c, v in enumerate(chunk):
            import pandas as pd
            names = list(fit.beta_samples.columns[:-2])
            df = pd.DataFrame(fit.chains[c], columns=names)
            df["iteration"] = fit.beta_samples["iteration"].to_numpy()
            df["burnin"] = fit.beta_samples["burnin"].to_numpy()
            st = dict(fit.stats["per_chain"][c])
            st["nEvaluations"] = st["ref_evals"]
            out.append(McmcGlm(beta_samples=df, beta_mean=df.loc[~df["burnin"], names].mean().to_frame().T, data=fit.data,
                               model_matrix=fit.model_matrix, param_list=None, family=fit.family, formula=fit.formula,
                               call=fit.call.replace(f"w = {chunk[0]}", f"w = {v}"), burnin=fit.burnin, sample_method=fit.sample_method,
                               qslice_fun=fit.qslice_fun, tuning={**fit.tuning, "w": v}, chains=fit.chains[c:c + 1], stats=st))
    return out
