"""Regenerates tests/golden/readme_gaussian.{npz,json}.

The reference prints, under seed 42, head(samples(norm)), coef(norm) and quantile(norm) for its
README example (reference README.md:73-120; data README.Rmd:44-55; seed README.Rmd:38).  R is not
installed in this image, so the inputs are regenerated with oracle/r_rng.py (R's Mersenne-Twister
+ inversion rnorm + rbinom restated) and the chain is replayed by the C oracle fed R's own runif
stream.  The JSON holds the numbers exactly as printed by the reference; the .npz holds the inputs
(X, y, beta0), the recorded uniform stream and the oracle's full-precision samples so that GPU
tests can replay the identical chain without R and without /root/reference.

Run from the repo root:  python tests/golden/make_readme_golden.py
"""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle.r_rng import RRng  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

README = {
    "source": "mathiaslj/mcmcglm README.md (rendered from README.Rmd with withr::local_seed(42))",
    "config": {"n": 1000, "formula": "Y ~ .", "family": "gaussian", "beta_prior": "dist_normal(0, 1)",
               "w": 0.5, "n_samples": 500, "burnin": 100, "sd": 1.0},
    # README.md:79-80 and :93-94
    "coef": {"(Intercept)": 1.011134, "X1": 1.490459, "X2": 2.026047},
    # README.md:103-106 (summarises burnin == TRUE rows: quirk Q2, R/mcmcglm_methods.R:137)
    "quantile": {
        "(Intercept)": {"mean": 1.014583, "q_0025": 0.8349995, "q_05": 1.013909, "q_0975": 1.099346},
        "X1": {"mean": 1.457432, "q_0025": 0.6940698, "q_05": 1.497910, "q_0975": 1.569321},
        "X2": {"mean": 1.996584, "q_0025": 1.8807500, "q_05": 2.024372, "q_0975": 2.177676},
    },
    # README.md:114-120
    "head_samples": [
        [0.6173367, -0.004541141, -0.09125636],
        [2.6508146, 0.281295470, 0.68343132],
        [0.8240996, 0.324627659, 2.30889073],
        [0.8170086, 1.028326905, 2.20351455],
        [0.8777350, 1.592074284, 2.16115289],
        [0.9092187, 1.442872350, 2.02913214],
    ],
    "head_burnin": [True] * 6,
}


def main():
    r = RRng(42)
    n = 1000
    x1 = r.rnorm(n)                       # README.Rmd:45
    x2 = r.rbinom_size1(n, 0.5)           # :46
    lin_pred = 1 + 1.5 * x1 + 2 * x2      # :47-50
    y = r.rnorm(n, lin_pred, 1.0)         # :52
    X = np.column_stack([np.ones(n), x1, x2])   # model.matrix(Y ~ .): (Intercept), X1, X2
    beta0 = r.rnorm(3)                    # R/mcmcglm.R:208 distributional::generate(dist_normal(0,1), 3)
    stream = r.runif(40000)               # what qslice's runif(1) calls will see
    m = oracle.make_model("gaussian", sd=1.0, prior="normal", prior_mu=0.0, prior_sigma=1.0)
    out = oracle.run_chain(m, X, y, beta0, w=0.5, n_iter=500, replay_u=stream)
    assert out["rc"] == 0
    used = out["uniforms_used"]
    samples = np.vstack([beta0, out["samples"]])
    np.savez_compressed(os.path.join(HERE, "readme_gaussian.npz"), X=X, y=y, beta0=beta0,
                        uniforms=stream[:used + 64], uniforms_used=np.int64(used), samples=samples,
                        n_eval=np.int64(out["n_eval"]), n_stepout=np.int64(out["n_stepout"]),
                        n_shrink=np.int64(out["n_shrink"]))
    with open(os.path.join(HERE, "readme_gaussian.json"), "w") as f:
        json.dump(README, f, indent=1)
    print("uniforms used", used, "evals", out["n_eval"], "stepouts", out["n_stepout"], "shrinks", out["n_shrink"])
    print(samples[:6])


if __name__ == "__main__":
    main()
