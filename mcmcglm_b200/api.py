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
        if self.kind == "gamma":                    # mu = shape, sigma = rate
            return rng.gamma(self.mu, 1.0 / self.sigma, n)
        if self.kind == "exponential":              # sigma = rate
            return rng.exponential(1.0 / self.sigma, n)
        return self.mu + self.sigma * rng.standard_t(self.df, n)

    def engine_tuple(self):
        return (self.kind, self.mu, self.sigma, self.df)


def dist_normal(mu=0.0, sigma=1.0):
    return Dist("normal", float(mu), float(sigma))


def dist_laplace(mu=0.0, sigma=1.0):
    return Dist("laplace", float(mu), float(sigma))


def dist_student_t(df, mu=0.0, sigma=1.0):
    return Dist("student_t", float(mu), float(sigma), float(df))


def dist_gamma(shape, rate):
    """distributional::dist_gamma(shape, rate) (vignettes/pospkg.Rmd:190-237): support (0, Inf), log-density -Inf outside."""
    return Dist("gamma", float(shape), float(rate))


def dist_exponential(rate):
    """distributional::dist_exponential(rate)"""
    return Dist("exponential", 0.0, float(rate))


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


def negative_binomial(theta=1.0, link="log"):
    """MASS::negative.binomial(theta) (vignettes/pospkg.Rmd:132-156).  Like the reference, whose log-density is
    dnbinom(Y, size = 1, mu = mu) whatever theta is (R/glm_utils.R:55-57), theta is accepted and ignored."""
    return Family("negative_binomial", link)


_FAMILY_BY_NAME = {"gaussian": gaussian, "binomial": binomial, "poisson": poisson, "negative_binomial": negative_binomial}
_LINKS_OF = {"gaussian": ("identity",), "binomial": ("logit", "probit"), "poisson": ("log",), "negative_binomial": ("log",)}


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


class ParamList(list):
    """param_list of the returned object: a list of {beta, eta, mu} with R-style (partly missing) names."""

    def __init__(self, items, names):
        super().__init__(items)
        self.names = names

    def __getitem__(self, k):
        if isinstance(k, str):
            return super().__getitem__(self.names.index(k))
        return super().__getitem__(k)


def _linkinv(fam, eta):
    """family$linkinv(eta) as stats computes it (R/mcmcglm.R:216, :269): only needed for param_list's `mu`, a value the
    slice path never reads (quirk Q8); evaluated on the host from the eta the engine returns."""
    eps = np.finfo(float).eps
    if fam.link == "identity":
        return eta.copy()
    if fam.link == "logit":
        e = np.where(eta < -30, eps, np.where(eta > 30, 1 / eps, np.exp(np.clip(eta, -30, 30))))
        return e / (1 + e)
    if fam.link == "probit":
        from scipy.special import ndtr
        return ndtr(np.clip(eta, -8.125890664701906, 8.125890664701906))
    with np.errstate(over="ignore"):
        return np.maximum(np.exp(eta), eps)


# ------------------------------------------------------------------------------ the front door
def _engine_kwargs(family, beta_prior, log_likelihood_extra_args):
    fam = check_family(family)
    if fam.family not in _LINKS_OF:
        raise L.CggError(L.E_UNSUPPORTED, f"family {fam.family!r} is not supported by the GPU engine (supported: {sorted(_LINKS_OF)})")
    if fam.link not in _LINKS_OF[fam.family]:
        raise L.CggError(L.E_UNSUPPORTED, f"link {fam.link!r} is not supported for family {fam.family!r} "
                         f"(supported: {_LINKS_OF[fam.family]})")
    # a list of priors: the reference sums EVERY prior at EVERY coordinate (R/glm_utils.R:113-115, quirk Q6) and draws
    # coordinate j of the start from prior j (R/mcmcglm.R:200-207)
    plist = list(beta_prior) if isinstance(beta_prior, (list, tuple)) else [beta_prior]
    for pr in plist:
        if not isinstance(pr, Dist):
            raise L.CggError(L.E_UNSUPPORTED, f"prior {pr!r} is not supported by the GPU engine "
                             "(supported: dist_normal, dist_laplace, dist_student_t, dist_gamma, dist_exponential and lists of them)")
    if len(plist) > L.MAX_PRIORS:
        raise L.CggError(L.E_UNSUPPORTED, f"a list of more than {L.MAX_PRIORS} priors is not supported by the GPU engine")
    sd = float((log_likelihood_extra_args or {}).get("sd", 1.0))
    p0 = plist[0]
    return fam, dict(family=fam.family, link=fam.link, sd=sd, prior=p0.kind, prior_mu=p0.mu, prior_sigma=p0.sigma, prior_df=p0.df,
                     more_priors=tuple(pr.engine_tuple() for pr in plist[1:]))


def mcmcglm(formula, family="gaussian", data=None, beta_prior=None, log_likelihood_extra_args=None,
            linear_predictor_calc="update", sample_method="slice_sampling", qslice_fun=slice_stepping_out,
            n_samples=500, burnin=100, *, n_chains=1, device=0, K=8, seed=None, beta_init=None,
            replay_uniforms=None, driver="persistent", keep_param_list=False, _w_per_chain=None, **tuning):
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
        if isinstance(beta_prior, (list, tuple)):                             # list: coordinate j from prior j, :200-207
            if len(beta_prior) != p:
                raise ValueError("The list length of the `beta_prior` specification needs to match the number of parameters "
                                 "in the model (potentially including intercept)")
            beta0 = np.stack([np.array([pr.generate(1, rng)[0] for pr in beta_prior]) for _ in range(n_chains)])
        else:
            beta0 = np.stack([beta_prior.generate(p, rng) for _ in range(n_chains)])
    else:
        beta0 = np.broadcast_to(np.asarray(beta_init, dtype=np.float64), (n_chains, p)).copy()
    mx = tuning.get("max", np.inf)
    eng_seed = int(rng.integers(0, 2 ** 63 - 1)) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 63))
    if np.isfinite(mx) and (mx != np.floor(mx)):
        raise L.CggError(L.E_ARG, "slice_stepping_out: a finite `max` must be a whole number")
    max_steps = -1 if np.isinf(mx) else max(int(mx), 0)                       # qslice: a finite max <= 0 means "no stepping out"
    param_list = None
    with Engine(n, p, w=float(tuning["w"]), max_steps=max_steps, n_chains=n_chains, K=K,
                device=device, driver=driver, seed=eng_seed, naive=(linear_predictor_calc == "naive"), **ekw) as e:
        e.set_data(X, Y)
        if _w_per_chain is not None:                                          # a tuning sweep: every chain its own w
            e.set_chain_w(_w_per_chain)
        for c in range(n_chains):
            e.init_chain(c, beta0[c])                                         # :215 init_eta = X %*% init_beta
        if not keep_param_list:
            S, st = e.run(n_samples, replay_u=replay_uniforms)                # :226-274
        else:
            # param_list (R/mcmcglm.R:183-189, :269, :287): beta, eta and mu of chain 1 after EVERY iteration -- (n_samples + 1)
            # n-vectors twice over (quirk Q4), so it is opt-in here; the engine is run one iteration at a time and read back
            if replay_uniforms is not None:
                raise L.CggError(L.E_ARG, "keep_param_list cannot be combined with replay_uniforms")
            b, eta = e.state(0)
            param_list = [dict(beta=b, eta=eta, mu=_linkinv(fam, eta))]
            parts, st = [], None
            for _ in range(n_samples):
                Sk, stk = e.run(1)
                parts.append(Sk)
                b, eta = e.state(0)
                param_list.append(dict(beta=b, eta=eta, mu=_linkinv(fam, eta)))
                st = stk if st is None else {k: (st[k] + v if isinstance(v, (int, float)) else [a + b_ for a, b_ in zip(st[k], v)] if k != "uniforms_used" else v)
                                             for k, v in stk.items()}
            S = np.concatenate(parts, axis=1)
        st["per_chain"] = [e.chain_stats(c) for c in range(n_chains)]
    import pandas as pd
    chains = np.concatenate([beta0[:, None, :], S], axis=1)                   # row 0 = the prior draw, :222
    df = pd.DataFrame(chains[0], columns=names)
    df["iteration"] = np.arange(n_samples + 1)
    df["burnin"] = df["iteration"] <= burnin + 1                              # :197-198 (quirk Q1)
    beta_mean = df.loc[~df["burnin"], names].mean().to_frame().T              # :276-280 (quirk Q3)
    pdesc = ", ".join(f"{q.kind}({q.mu:g}, {q.sigma:g})" for q in (beta_prior if isinstance(beta_prior, (list, tuple)) else [beta_prior]))
    call = (f"mcmcglm(formula = {formula}, family = \"{fam.family}\", data = <data>, beta_prior = {pdesc}, "
            + ", ".join(f"{k} = {v}" for k, v in tuning.items()) + ")")
    if param_list is not None:
        # names as the reference builds them (R/mcmcglm.R:183-189): `ifelse` with a scalar test keeps only "burnin1", so the
        # names run out before the list does and the remaining entries are unnamed (NA) -- quirk Q4
        nm = ["init"] + (["burnin1"] if burnin != 0 else []) + [f"iteration{i}" for i in range(1, n_samples - burnin + 1)]
        nm = (nm + [None] * (n_samples + 1))[:n_samples + 1]
        param_list = ParamList(param_list, nm)
    return McmcGlm(beta_samples=df, beta_mean=beta_mean, data=data, model_matrix=X, param_list=param_list, family=fam,
                   formula=formula, call=call, burnin=burnin, sample_method=sample_method, qslice_fun=qslice_fun,
                   tuning=dict(tuning), chains=chains, stats=st)


# ------------------------------------------------------------------------------ operators
def log_potential_from_betaj(new_beta_j, j, current_beta, current_eta, Y, X, family, beta_prior,
                             linear_predictor_calc="update", device=0, **extra):
    """log_potential_from_betaj of R/glm_utils.R:187-218 evaluated by the GPU kernel (K1).
    `j` is 1-based, as in the reference.  `new_beta_j` may be a scalar or an array of candidates."""
    if linear_predictor_calc not in ("update", "naive"):
        raise ValueError("'arg' should be one of 'update', 'naive'")
    _, ekw = _engine_kwargs(family, beta_prior, {"sd": extra.get("sd", 1.0)})
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    if not 1 <= j <= p:
        raise IndexError("j is 1-based and must be in 1..ncol(X)")
    with Engine(n, p, w=1.0, n_chains=1, device=device, driver="stepwise", **ekw) as e:
        e.set_data(X, Y)
        if linear_predictor_calc == "naive":     # R/glm_utils.R:206-208: new_eta <- X %*% new_beta; current_eta is not used
            e.init_chain(0, current_beta)        # eta = X %*% current_beta on the device (K4), then eta + X_j (b - beta_j)
        else:
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


# ------------------------------------------------------------------------------ runtime comparison (R/measure_performance.R)
def generate_normal_data(n_vars, n=100, beta=None, sd=1.0, rng=None):
    """generate_normal_data of R/measure_performance.R:46-63: gaussian response, intercept + (n_vars - 1) N(0, 1) columns."""
    import pandas as pd
    rng = np.random.default_rng() if rng is None else rng
    beta = np.ones(n_vars) if beta is None else np.asarray(beta, dtype=np.float64)
    Xs = rng.standard_normal((n, n_vars - 1))
    y = beta[0] + Xs @ beta[1:] + sd * rng.standard_normal(n)
    d = pd.DataFrame(Xs, columns=[f"X{i}" for i in range(1, n_vars)])
    d.insert(0, "Y", y)
    return d


def compare_eta_comptime(formula, family="gaussian", data=None, beta_prior=None, log_likelihood_extra_args=None,
                         sample_method="slice_sampling", qslice_fun=slice_stepping_out, n_samples=500, burnin=100, **tuning):
    """compare_eta_comptime of R/measure_performance.R:3-42: the same mcmcglm() call timed with linear_predictor_calc =
    "update" (the O(n) CGGibbs update) and "naive" (eta recomputed from scratch, O(n p)); one row per method."""
    import time
    import pandas as pd
    beta_prior = dist_normal(0, 1) if beta_prior is None else beta_prior
    lla = {"sd": 1} if log_likelihood_extra_args is None else log_likelihood_extra_args
    rows = []
    for calc in ("update", "naive"):
        t0 = time.perf_counter()
        m = mcmcglm(formula, family, data, beta_prior, lla, calc, sample_method, qslice_fun, n_samples, burnin,
                    driver="stepwise", **tuning)
        rows.append({"time": time.perf_counter() - t0, "linear_predictor_calc": calc, "n_vars": m.model_matrix.shape[1],
                     "n_samples": n_samples, "beta_mean": getattr(beta_prior, "mu", None), "beta_variance": getattr(beta_prior, "sigma", 1.0) ** 2,
                     "family": m.family.family, **lla, "qslice_fun": "qslice::slice_stepping_out", **tuning})
    return pd.DataFrame(rows)


def compare_eta_comptime_across_nvars(n_vars, n=100, beta_prior=None, log_likelihood_extra_args=None, sample_method="slice_sampling",
                                      qslice_fun=slice_stepping_out, n_samples=500, burnin=100, parallelise=False, n_cores=None,
                                      rng=None, **tuning):
    """compare_eta_comptime_across_nvars of R/measure_performance.R:113-151 (the experiment of vignettes/performance.Rmd:31-41):
    for every number of variables, simulate gaussian data and time "update" against "naive".  w defaults to 0.5 like the
    reference's; parallelise / n_cores are accepted and ignored (one GPU)."""
    import pandas as pd
    if qslice_fun is slice_stepping_out and not tuning:
        tuning = {"w": 0.5}
    out = []
    for nv in np.atleast_1d(n_vars):
        d = generate_normal_data(int(nv), n=n, rng=rng)
        out.append(compare_eta_comptime("Y ~ .", "gaussian", d, beta_prior, log_likelihood_extra_args, sample_method, qslice_fun,
                                        n_samples, burnin, **tuning))
    res = pd.concat(out, ignore_index=True)
    res["parallelised"] = bool(parallelise)
    return res
