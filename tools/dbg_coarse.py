import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine
X, y, bt = synth("binomial", 20001, 6, seed=31)
beta0 = np.random.default_rng(1).standard_normal((3, 6))
def run(pref, iters=30, C=3, theta=None):
    if theta is not None: os.environ["CGG_COARSE_THETA"]=str(theta)
    else: os.environ.pop("CGG_COARSE_THETA",None)
    with Engine(20001, 6, family="binomial", w=0.5, n_chains=C, K=8, seed=4, prefilter=pref, **PRIOR_CASES["laplace"]) as e:
        e.set_data(X, y)
        for c in range(C): e.init_chain(c, beta0[c])
        return e.run(iters)
S0, st0 = run(False); S1, st1 = run(True)
print({k:st0[k] for k in ('cand_evals','coarse_evals','coarse_undecided','ref_evals','passes','uniforms_used')})
print({k:st1[k] for k in ('cand_evals','coarse_evals','coarse_undecided','ref_evals','passes','uniforms_used')})
d = np.argwhere(S0 != S1)
print("n diffs", len(d), "first", d[:3])
if len(d):
    c,it,j = d[0]
    print("chain",c,"iter",it,"j",j, S0[c,it,j], S1[c,it,j])
