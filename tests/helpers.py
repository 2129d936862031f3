"""Shared synthetic-data helpers for tests and bench (SURVEY.md 8(d) recipe)."""
import numpy as np


def synth(family, n, p, seed=0, intercept=True):
    """X: column 0 == 1, others iid N(0,1); beta* ~ N(0, 1/p) so sd(eta*) ~ 1; y drawn from the family."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    if intercept:
        X[:, 0] = 1.0
    X = np.asfortranarray(X)
    beta_true = rng.standard_normal(p) / np.sqrt(p)
    eta = X @ beta_true
    if family == "gaussian":
        y = eta + rng.standard_normal(n)
    elif family == "binomial":
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    elif family == "poisson":
        y = rng.poisson(np.exp(eta)).astype(np.float64)
    else:
        raise ValueError(family)
    return X, y, beta_true


PRIOR_CASES = {
    "normal": dict(prior="normal", prior_mu=0.0, prior_sigma=1.0),
    "laplace": dict(prior="laplace", prior_mu=0.0, prior_sigma=1.0),
    "student_t": dict(prior="student_t", prior_mu=0.0, prior_sigma=1.0, prior_df=4.0),
}
