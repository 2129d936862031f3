"""Engine: a thin object wrapper over the C ABI (include/cggibbs.h).  All compute happens in
libcggibbs.so on the GPU; this file only marshals numpy buffers."""
import ctypes as C
import numpy as np
from . import _lib as L

FAMILIES = {"gaussian": (L.GAUSSIAN, L.LINK_IDENTITY), "binomial": (L.BINOMIAL, L.LINK_LOGIT),
            "poisson": (L.POISSON, L.LINK_LOG), "negative_binomial": (L.NEGATIVE_BINOMIAL, L.LINK_LOG)}
LINKS = {"identity": L.LINK_IDENTITY, "logit": L.LINK_LOGIT, "log": L.LINK_LOG, "probit": L.LINK_PROBIT}
PRIORS = {"normal": L.PRIOR_NORMAL, "laplace": L.PRIOR_LAPLACE, "student_t": L.PRIOR_STUDENT_T, "gamma": L.PRIOR_GAMMA,
          "exponential": L.PRIOR_EXPONENTIAL}
# kernel-side family codes (cgg_debug_row_terms): binomial with the probit link is a family of its own there
ROW_TERM_FAMILIES = {"gaussian": 0, "binomial": 1, "poisson": 2, "negative_binomial": 3, "binomial_probit": 4}
# "persistent": the engine picks the grid-wide kernel or, for small n, one cluster per chain; "grid" / "cluster" force one
DRIVERS = {"persistent": L.DRIVER_PERSISTENT, "grid": L.DRIVER_PERSISTENT, "cluster": L.DRIVER_CLUSTER, "stepwise": L.DRIVER_STEPWISE}

_dp = C.POINTER(C.c_double)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def debug_row_terms(family, y, eta, sd=1.0, device=0):
    """Per-row log-density terms exactly as the kernels accumulate them (diagnostic; see cggibbs.h)."""
    lib = L.load()
    y, eta = _f64(y).ravel(), _f64(eta).ravel()
    out = np.empty_like(y)
    L.check(lib.cgg_debug_row_terms(device, ROW_TERM_FAMILIES[family], y.size, y.ctypes.data_as(_dp), eta.ctypes.data_as(_dp),
                                    float(sd), out.ctypes.data_as(_dp)))
    return out


class Engine:
    """One device's CGGibbs state: data (X, y) + n_chains chains (beta, eta, slice state, RNG)."""

    def __init__(self, n, p, family="gaussian", link=None, sd=1.0, prior="normal", prior_mu=0.0, prior_sigma=1.0,
                 prior_df=1.0, w=0.5, max_steps=-1, n_chains=1, K=8, device=0, driver="persistent", seed=0,
                 chain_offset=0, spec_tau=0.12, rows_per_cta_min=0, row_sharded=False, prefilter=True, jet=True,
                 jet_light=True, jet_bound_scale=1.0, more_priors=(), naive=False):
        """more_priors: further components of a LIST of priors (every component is evaluated at every coordinate, like the
        reference does): tuples (kind, a, b, c) = (mu, sigma, df) | gamma (shape, rate, -) | exponential (-, rate, -)."""
        self._h = None
        self._lib = L.load()
        if family not in FAMILIES:
            raise L.CggError(L.E_UNSUPPORTED, f"unsupported family {family!r}; supported: {sorted(FAMILIES)}")
        fam, canon = FAMILIES[family]
        if link is None:
            lnk = canon
        elif link in LINKS:
            lnk = LINKS[link]
        else:
            raise L.CggError(L.E_UNSUPPORTED, f"unsupported link {link!r}; supported: {sorted(LINKS)}")
        if prior not in PRIORS:
            raise L.CggError(L.E_UNSUPPORTED, f"unsupported prior {prior!r}; supported: {sorted(PRIORS)}")
        if max_steps is None or (isinstance(max_steps, float) and np.isinf(max_steps)):
            max_steps = -1
        cfg = L.Config(abi_version=L.ABI_VERSION, device=device, n=n, p=p, family=fam, link=lnk, sd=sd,
                       prior=PRIORS[prior], n_chains=n_chains, prior_mu=prior_mu, prior_sigma=prior_sigma,
                       prior_df=prior_df, w=w, max_steps=int(max_steps), K=K, driver=DRIVERS[driver],
                       mode=L.MODE_ROW_SHARDED if row_sharded else L.MODE_CHAINS, chain_offset=chain_offset,
                       seed=seed, spec_tau=spec_tau, rows_per_cta_min=rows_per_cta_min,
                       flags=((0 if prefilter else L.FLAG_NO_PREFILTER) | (0 if jet else L.FLAG_NO_JET)
                              | (0 if jet_light else L.FLAG_NO_JET_LIGHT) | (L.FLAG_NO_CLUSTER if driver == "grid" else 0)
                              | (L.FLAG_NAIVE if naive else 0)),
                       jet_bound_scale=jet_bound_scale)
        h = C.c_void_p()
        L.check(self._lib.cgg_create(C.byref(cfg), C.byref(h)))
        self._h = h
        for kind, a, b, c in more_priors:
            if kind not in PRIORS:
                raise L.CggError(L.E_UNSUPPORTED, f"unsupported prior {kind!r}; supported: {sorted(PRIORS)}")
            L.check(self._lib.cgg_add_prior(self._h, PRIORS[kind], float(a), float(b), float(c)))
        self.n, self.p, self.n_chains = n, p, n_chains
        self._keep = []

    def close(self):
        if self._h is not None:
            self._lib.cgg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- data ------------------------------------------------------------------------------
    def set_data(self, X, y):
        """X: n x p (any layout; sent column-major like an R matrix), y: n.  Host buffers."""
        X = np.asarray(X, dtype=np.float64)
        if X.shape != (self.n, self.p):
            raise L.CggError(L.E_ARG, f"X has shape {X.shape}, expected {(self.n, self.p)}")
        Xf = np.asfortranarray(X)
        y = _f64(y)
        if y.shape != (self.n,):
            raise L.CggError(L.E_ARG, f"y has shape {y.shape}, expected {(self.n,)}")
        L.check(self._lib.cgg_set_data(self._h, Xf.ctypes.data, self.n, y.ctypes.data))

    def set_data_ptr(self, X_ptr, ldx, y_ptr, device=False, keepalive=None):
        """Raw-pointer variant (pinned host buffers or device memory of torch tensors)."""
        fn = self._lib.cgg_set_data_device if device else self._lib.cgg_set_data
        L.check(fn(self._h, X_ptr, ldx, y_ptr))
        self._keep = [keepalive]

    # -- chains ----------------------------------------------------------------------------
    def init_chain(self, chain, beta0):
        b = _f64(beta0)
        if b.shape != (self.p,):
            raise L.CggError(L.E_ARG, f"beta0 has shape {b.shape}, expected {(self.p,)}")
        L.check(self._lib.cgg_init_chain(self._h, chain, b.ctypes.data_as(_dp)))

    def set_state(self, chain, beta, eta):
        b, e = _f64(beta), _f64(eta)
        if b.shape != (self.p,) or e.shape != (self.n,):
            raise L.CggError(L.E_ARG, "set_state: beta must have p entries and eta n entries")
        L.check(self._lib.cgg_set_state(self._h, chain, b.ctypes.data_as(_dp), e.ctypes.data_as(_dp)))

    def log_potential(self, chain, j, cands):
        c = np.atleast_1d(_f64(cands))
        out = np.empty_like(c)
        L.check(self._lib.cgg_log_potential(self._h, chain, j, c.size, c.ctypes.data_as(_dp), out.ctypes.data_as(_dp)))
        return out

    def debug_jet(self, chain, j, cands, light=False):
        """One jet pass along column j (no state change): (surrogate log-likelihood, error bound, raw sums);
        light=True (binomial): the log-likelihood difference to the current point instead."""
        c = np.atleast_1d(_f64(cands))
        val, bnd, sums = np.empty_like(c), np.empty_like(c), np.empty(L.JET_NV)
        L.check(self._lib.cgg_debug_jet(self._h, chain, j, c.size, int(bool(light)), c.ctypes.data_as(_dp), val.ctypes.data_as(_dp),
                                        bnd.ctypes.data_as(_dp), sums.ctypes.data_as(_dp)))
        return val, bnd, sums

    def update_eta(self, chain, j, new_beta_j):
        L.check(self._lib.cgg_update_eta(self._h, chain, j, float(new_beta_j)))

    def state(self, chain, want_eta=True):
        beta = np.empty(self.p)
        eta = np.empty(self.n) if want_eta else None
        L.check(self._lib.cgg_get_state(self._h, chain, beta.ctypes.data_as(_dp),
                                        eta.ctypes.data_as(_dp) if want_eta else None))
        return beta, eta

    def fx(self, chain):
        v = C.c_double()
        L.check(self._lib.cgg_get_fx(self._h, chain, C.byref(v)))
        return v.value

    def run(self, n_iter, replay_u=None, want_samples=True):
        """Runs n_iter Gibbs iterations on every chain.  Returns (samples[C, n_iter, p], stats dict)."""
        st = L.Stats()
        used = (C.c_uint64 * self.n_chains)()
        samples = np.empty((self.n_chains, n_iter, self.p)) if want_samples else None
        if replay_u is not None:
            ru = _f64(replay_u)
            if ru.ndim == 1:
                ru = ru[None, :]
            if ru.shape[0] != self.n_chains:
                raise L.CggError(L.E_ARG, "replay_u needs one row per chain")
            ru = np.ascontiguousarray(ru)
            ru_p, n_u = ru.ctypes.data, ru.shape[1]
        else:
            ru_p, n_u = None, 0
        rc = self._lib.cgg_run(self._h, n_iter, ru_p, n_u, used, samples.ctypes.data if want_samples else None,
                               C.byref(st))
        d = st.as_dict()
        d["uniforms_used"] = list(used)
        self.last_stats = d
        L.check(rc)
        return samples, d

    def set_chain_w(self, w):
        """Per-chain slice widths (one engine run serves a whole tuning sweep); w: n_chains values."""
        w = _f64(w)
        if w.shape != (self.n_chains,):
            raise L.CggError(L.E_ARG, f"w has shape {w.shape}, expected {(self.n_chains,)}")
        L.check(self._lib.cgg_set_chain_w(self._h, w.ctypes.data_as(_dp)))

    def chain_stats(self, chain):
        """Counters of one chain from the last run (ref_evals = qslice's nEvaluations)."""
        st = L.Stats()
        L.check(self._lib.cgg_get_chain_stats(self._h, chain, C.byref(st)))
        return st.as_dict()

    def launch_shape(self):
        a, b = C.c_int32(), C.c_int32()
        L.check(self._lib.cgg_launch_shape(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream_ptr(self):
        return self._lib.cgg_stream(self._h)

    def comm_init_nccl(self, rank, world, id_bytes):
        """Row-sharded mode: join the NCCL communicator identified by the 128-byte id (see multigpu.init_nccl)."""
        L.check(self._lib.cgg_comm_init_nccl(self._h, rank, world, id_bytes))

    def p2p_mailbox(self, world):
        """Row-sharded + persistent driver: allocates this rank's mailbox; returns (device pointer, 64-byte IPC handle)."""
        ptr = C.c_void_p()
        buf = C.create_string_buffer(64)
        L.check(self._lib.cgg_p2p_mailbox(self._h, world, C.byref(ptr), buf))
        return ptr.value, buf.raw

    def p2p_connect(self, rank, world, dev_ptrs=None, ipc_handles=None):
        """Maps the peers' mailboxes: dev_ptrs (same process) or the concatenated IPC handles (other processes)."""
        arr = None
        if dev_ptrs is not None:
            arr = (C.c_void_p * world)(*[C.c_void_p(p) if p else None for p in dev_ptrs])
        L.check(self._lib.cgg_p2p_connect(self._h, rank, world, arr, b"".join(ipc_handles) if ipc_handles is not None else None))

    def set_exchange(self, fn):
        """fn(device_ptr:int, count:int, stream_ptr:int) -> int; kept alive by the engine."""
        def tramp(user, buf, count, stream):
            try:
                return int(fn(buf, count, stream) or 0)
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._xfn = L.EXCHANGE_FN(tramp)
        L.check(self._lib.cgg_set_exchange(self._h, self._xfn, None))
