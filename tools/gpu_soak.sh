# long runs through many launches: no hangs, finite samples, posterior sanity
mkdir -p gpurun_out
timeout 900 python - <<'PY' > gpurun_out/soak.log 2>&1
import numpy as np, time, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import synth
from mcmcglm_b200 import Engine
for family, prior, n, p, C, iters in (("binomial", "laplace", 100000, 100, 4, 1500), ("poisson", "student_t", 200000, 50, 8, 800), ("gaussian", "normal", 300000, 40, 6, 1200), ("binomial", "normal", 5000, 8, 32, 3000)):
    X, y, bt = synth(family, n, p, seed=7)
    kw = dict(prior=prior, prior_mu=0.0, prior_sigma=1.0, prior_df=4.0)
    with Engine(n, p, family=family, w=0.5, n_chains=C, seed=3, **kw) as e:
        e.set_data(X, y)
        rng = np.random.default_rng(1)
        for c in range(C): e.init_chain(c, 0.3 * rng.standard_normal(p))
        t = time.perf_counter(); S, st = e.run(iters); dt = time.perf_counter() - t
    m = S[:, iters // 3:, :].mean(axis=(0, 1))
    print(family, n, p, C, iters, "launches", st["launches"], "updates/s %.0f" % (st["updates"] / dt), "finite", bool(np.isfinite(S).all()),
          "max |mean - truth| %.3g" % np.max(np.abs(m - bt)), "fallbacks", st["jet_fallbacks"], "retries", st["jet_retries"], flush=True)
PY
cat gpurun_out/soak.log
