"""ctypes binding of include/cggibbs.h.  There is no fallback: if libcggibbs.so is missing or does
not load, importing the engine raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libcggibbs.so")

ABI_VERSION = 3
KMAX = 8
OK, E_ARG, E_UNSUPPORTED, E_CUDA, E_NAN, E_STREAM, E_NOTERM, E_STATE, E_COMM = 0, -1, -2, -3, -4, -5, -6, -7, -8
GAUSSIAN, BINOMIAL, POISSON, NEGATIVE_BINOMIAL = 0, 1, 2, 3
LINK_IDENTITY, LINK_LOGIT, LINK_LOG, LINK_PROBIT = 0, 1, 2, 3
PRIOR_NORMAL, PRIOR_LAPLACE, PRIOR_STUDENT_T, PRIOR_GAMMA, PRIOR_EXPONENTIAL = 0, 1, 2, 3, 4
MAX_PRIORS = 8
DRIVER_PERSISTENT, DRIVER_STEPWISE, DRIVER_CLUSTER = 0, 1, 2
MODE_CHAINS, MODE_ROW_SHARDED = 0, 1
FLAG_NO_PREFILTER = 1
FLAG_NO_JET = 2
FLAG_NO_JET_LIGHT = 4
FLAG_NO_CLUSTER = 8
FLAG_NAIVE = 16
JET_NV = KMAX + 2

# every symbol include/cggibbs.h declares
EXPORTS = ["cgg_last_error", "cgg_abi_version", "cgg_create", "cgg_destroy", "cgg_set_data",
           "cgg_set_data_device", "cgg_init_chain", "cgg_set_state", "cgg_log_potential", "cgg_update_eta", "cgg_run",
           "cgg_get_state", "cgg_get_fx", "cgg_set_exchange", "cgg_stream", "cgg_launch_shape", "cgg_debug_row_terms",
           "cgg_nccl_unique_id", "cgg_comm_init_nccl", "cgg_debug_coarse_error", "cgg_debug_jet", "cgg_debug_light_error", "cgg_set_chain_w",
           "cgg_get_chain_stats", "cgg_p2p_mailbox", "cgg_p2p_connect", "cgg_add_prior"]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n", C.c_int64), ("p", C.c_int64),
                ("family", C.c_int32), ("link", C.c_int32), ("sd", C.c_double), ("prior", C.c_int32),
                ("n_chains", C.c_int32), ("prior_mu", C.c_double), ("prior_sigma", C.c_double),
                ("prior_df", C.c_double), ("w", C.c_double), ("max_steps", C.c_int64), ("K", C.c_int32),
                ("driver", C.c_int32), ("mode", C.c_int32), ("chain_offset", C.c_int32), ("seed", C.c_uint64),
                ("spec_tau", C.c_double), ("rows_per_cta_min", C.c_int32), ("flags", C.c_int32),
                ("jet_bound_scale", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("updates", C.c_uint64), ("passes", C.c_uint64), ("chain_passes", C.c_uint64),
                ("commit_passes", C.c_uint64), ("cand_evals", C.c_uint64), ("ref_evals", C.c_uint64),
                ("stepouts", C.c_uint64), ("shrinks", C.c_uint64), ("launches", C.c_uint64),
                ("sweep_ms", C.c_double), ("algorithmic_bytes", C.c_double), ("coarse_evals", C.c_uint64),
                ("coarse_undecided", C.c_uint64), ("jet_passes", C.c_uint64), ("jet_fallbacks", C.c_uint64),
                ("jet_retries", C.c_uint64), ("group_passes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

_lib = None


class CggError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cgg {code}] {msg}")
        self.code = code


def load():
    """Loads libcggibbs.so (building is the job of __graft_entry__.build / mcmcglm_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CGG_LIB", SO_PATH)      # experiments only: an alternative build of the same ABI
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m mcmcglm_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, dp, i32, i64, u64 = C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_int64, C.c_uint64
    L.cgg_last_error.restype = C.c_char_p
    L.cgg_last_error.argtypes = []
    L.cgg_abi_version.restype = C.c_int
    L.cgg_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.cgg_destroy.argtypes = [vp]
    L.cgg_destroy.restype = None
    L.cgg_set_data.argtypes = [vp, vp, i64, vp]
    L.cgg_set_data_device.argtypes = [vp, vp, i64, vp]
    L.cgg_init_chain.argtypes = [vp, i32, dp]
    L.cgg_set_state.argtypes = [vp, i32, dp, dp]
    L.cgg_log_potential.argtypes = [vp, i32, i64, i32, dp, dp]
    L.cgg_update_eta.argtypes = [vp, i32, i64, C.c_double]
    L.cgg_run.argtypes = [vp, i64, vp, u64, C.POINTER(u64), vp, C.POINTER(Stats)]
    L.cgg_get_state.argtypes = [vp, i32, dp, dp]
    L.cgg_get_fx.argtypes = [vp, i32, dp]
    L.cgg_set_exchange.argtypes = [vp, EXCHANGE_FN, vp]
    L.cgg_nccl_unique_id.argtypes = [C.c_char_p]
    L.cgg_comm_init_nccl.argtypes = [vp, i32, i32, C.c_char_p]
    L.cgg_debug_coarse_error.argtypes = [i32, dp, dp]
    L.cgg_debug_jet.argtypes = [vp, i32, i64, i32, i32, dp, dp, dp, dp]
    L.cgg_debug_light_error.argtypes = [i32, dp]
    L.cgg_set_chain_w.argtypes = [vp, dp]
    L.cgg_add_prior.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double]
    L.cgg_p2p_mailbox.argtypes = [vp, i32, C.POINTER(vp), C.c_char_p]
    L.cgg_p2p_connect.argtypes = [vp, i32, i32, C.POINTER(vp), C.c_char_p]
    L.cgg_get_chain_stats.argtypes = [vp, i32, C.POINTER(Stats)]
    L.cgg_debug_row_terms.argtypes = [i32, i32, i64, dp, dp, C.c_double, dp]
    L.cgg_stream.argtypes = [vp]
    L.cgg_stream.restype = vp
    L.cgg_launch_shape.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    if L.cgg_abi_version() != ABI_VERSION:
        raise ImportError("libcggibbs.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise CggError(rc, load().cgg_last_error().decode())
