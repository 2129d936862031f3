"""How a chain started from a prior draw (the reference's start, R/mcmcglm.R:200-222) reaches the stationary regime on the
headline workload: per Gibbs iteration, time, passes per update, candidate evaluations, pre-filter share, jet passes."""
import argparse
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mcmcglm_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--cols", type=int)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
if a.cols:
    wl["p"] = a.cols
dev = torch.device("cuda", 0)
X, y = bench.make_data(wl, dev, 42)
n, p, C = wl["n"], wl["p"], wl["chains"]
beta0 = bench.draw_beta0(wl, np.random.default_rng(42), C)
e = Engine(n, p, family=wl["family"], w=wl["w"], n_chains=C, K=wl["K"], seed=42, spec_tau=0.12, **bench.PRIOR_KW[wl["prior"]])
e.set_data_ptr(X.data_ptr(), n, y.data_ptr(), device=True, keepalive=(X, y))
for c in range(C):
    e.init_chain(c, beta0[c])
tot = 0.0
for it in range(a.iters):
    t0 = time.perf_counter()
    _, st = e.run(1, want_samples=False)
    dt = time.perf_counter() - t0
    tot += dt
    u = max(st["updates"], 1)
    sd_eta = float(np.std(e.state(0)[1])) if it % 5 == 0 else float("nan")
    print(json.dumps({"iter": it + 1, "ms": round(1e3 * dt, 1), "kernel_ms": round(st["sweep_ms"], 1), "passes_per_update": round(st["chain_passes"] / u, 2),
                      "cand_evals_per_update": round(st["cand_evals"] / u, 2), "coarse_share": round(st["coarse_evals"] / max(st["cand_evals"], 1), 2),
                      "coarse_undecided_per_update": round(st["coarse_undecided"] / u, 3), "jet_passes_per_update": round(st["jet_passes"] / u, 2),
                      "jet_fallbacks_per_update": round(st["jet_fallbacks"] / u, 3), "ref_evals_per_update": round(st["ref_evals"] / u, 2),
                      "stepouts_per_update": round(st["stepouts"] / u, 2), "sd_eta_chain0": round(sd_eta, 2)}))
print(json.dumps({"total_s": round(tot, 2)}))
