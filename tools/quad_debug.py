"""Diagnostic: a short run with group passes on a mid-sized problem; prints the stats or the error."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine

family = sys.argv[1] if len(sys.argv) > 1 else "binomial"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150001
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
p = 5
X, y, bt = synth(family, n, p, seed=29)
beta0 = bt + 0.05 * np.random.default_rng(5).standard_normal((C, p))
with Engine(n, p, family=family, sd=1.0, n_chains=C, w=0.5, driver="grid", **PRIOR_CASES["laplace"]) as e:
    e.set_data(X, y)
    for c in range(C):
        e.init_chain(c, beta0[c])
    S, st = e.run(iters)
    print({k: st[k] for k in ("updates", "passes", "jet_passes", "jet_fallbacks", "jet_retries", "group_passes", "launches")}, float(S.sum()))
