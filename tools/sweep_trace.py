"""Per-sweep trace of the engine on a bench workload: ms, evaluations per update, pre-filter share.
   python tools/sweep_trace.py --cols 100 --sweeps 60 [--scale 1.0] [--theta 1.0]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from mcmcglm_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=100); ap.add_argument("--sweeps", type=int, default=60)
ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--chains", type=int, default=8)
ap.add_argument("--every", type=int, default=5); ap.add_argument("--nopre", action="store_true")
ap.add_argument("--tau", type=float, default=0.12); ap.add_argument("--K", type=int, default=8)
a = ap.parse_args()
wl = dict(bench.WORKLOADS["cfg3"]); wl["p"] = a.cols; wl["chains"] = a.chains
dev = torch.device("cuda", 0)
X, y = bench.make_data(wl, dev, 42)
rng = np.random.default_rng(42)
beta0 = bench.draw_beta0(wl, rng, a.chains) * a.scale
e = Engine(wl["n"], wl["p"], family="binomial", w=0.5, n_chains=a.chains, K=a.K, seed=42, prefilter=not a.nopre, spec_tau=a.tau, **bench.PRIOR_KW["laplace"])
e.set_data_ptr(X.data_ptr(), wl["n"], y.data_ptr(), device=True, keepalive=(X, y))
for c in range(a.chains):
    e.init_chain(c, beta0[c])
for s in range(a.sweeps):
    _, st = e.run(1, want_samples=False)
    if s % a.every == 0 or s == a.sweeps - 1:
        u = st["updates"]
        print(f"sweep {s:3d}: {st['sweep_ms']:8.2f} ms  ref_evals/upd {st['ref_evals']/u:5.2f}  cand/upd {st['cand_evals']/u:5.2f}  "
              f"chainpass/upd {st['chain_passes']/u:4.2f}  stepouts/upd {st['stepouts']/u:4.2f} shrinks/upd {st['shrinks']/u:4.2f}  "
              f"prefiltered {st['coarse_evals']/max(st['cand_evals'],1):5.3f} undecided/upd {st['coarse_undecided']/u:5.3f}  upd/s {u/st['sweep_ms']*1e3:9.0f}")
