/*
 * cggibbs.h -- C ABI of the B200-native CGGibbs engine (libcggibbs.so).
 *
 * This is the drop-in boundary for ONE path of mathiaslj/mcmcglm: the slice-within-Gibbs
 * coordinate update (reference R/mcmcglm.R:226-274 and everything it calls).  The reference is pure
 * R and has no FFI of its own; these entry points are what an R `.Call` shim (r/src/rshim.c), a
 * ctypes binding (mcmcglm_b200/_lib.py) or any other host binds.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns CGG_OK (0) or a negative cgg_status; cgg_last_error() gives the
 *     message of the calling thread's last failure.  Nothing here exits, aborts or prints.
 *   - "host" pointers are ordinary CPU memory owned by the caller; the library copies in/out.
 *     "device" pointers (only cgg_set_data_device) are CUDA device memory on cfg.device.
 *   - matrices are column-major fp64 exactly as R stores them: X[i + j*ldx], ldx >= n.
 *   - j is 0-based here; the R shim subtracts 1.
 *   - a handle is not re-entrant; all calls block until their results are in the output buffers.
 *   - environment switches read by the library (experiments and tests only; none of them can change a result):
 *     CGG_PAIR=0|1 (pair passes off/on; default on from 4 chains), CGG_CHUNK=<iterations per launch when pair passes
 *     are on; default 8>, CGG_COLCACHE=0|1 (per-warp X-column cache of the pair passes), CGG_EARLY=0|1 (the plain update
 *     judged, published and booked straight from the deciding warp's cache; default on), CGG_SMALLN=0|1 (cluster driver
 *     for n <= 2^18), CGG_L2_PERSIST=0|1, CGG_QUAD=0|1 (group passes, only in builds with -DCGG_GROUP_PASSES),
 *     CGG_COARSE_THETA=<fp32 pre-filter policy>, CGG_PROFILE[_TRACE|_CTAS|_WARPS]=1 (phase counters of the persistent
 *     kernel on stderr), CGG_DEBUG_PTRS=1 (device addresses of the handle's buffers on stderr).
 *   - there is NO CPU fallback: unsupported family/link/prior/sampler => CGG_E_UNSUPPORTED,
 *     no usable CUDA device => CGG_E_CUDA.
 */
#ifndef CGGIBBS_H
#define CGGIBBS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGG_ABI_VERSION 3
#define CGG_KMAX 8 /* most candidates one chain can score in one pass over its rows */

typedef enum cgg_status {
    CGG_OK = 0,
    CGG_E_ARG = -1,         /* bad argument (NULL, size, alignment, y outside the family's support) */
    CGG_E_UNSUPPORTED = -2, /* family/link/prior/sampler outside the supported set: rejected */
    CGG_E_CUDA = -3,        /* CUDA runtime failure, or no device */
    CGG_E_NAN = -4,         /* log-potential evaluated to NaN where the reference would stop() */
    CGG_E_STREAM = -5,      /* replay-uniform stream exhausted */
    CGG_E_NOTERM = -6,      /* slice loop exceeded the pass guard (reference would spin forever) */
    CGG_E_STATE = -7,       /* call sequence error (no data, chain not initialised, ...) */
    CGG_E_COMM = -8         /* collective layer failure (row-sharded mode) */
} cgg_status;

/* family$family / family$link of the reference (R/family_data_processing.R:3-16).  Implemented pairs: gaussian+identity,
 * binomial+logit, poisson+log (all drivers, jet passes) and -- exact passes on the stepwise driver -- binomial+probit
 * (vignettes/pospkg.Rmd:88-108) and negative binomial+log (MASS::negative.binomial; the reference's log-density is
 * dnbinom(size = 1) whatever theta is, R/glm_utils.R:55-57). */
enum { CGG_GAUSSIAN = 0, CGG_BINOMIAL = 1, CGG_POISSON = 2, CGG_NEGATIVE_BINOMIAL = 3 };
enum { CGG_LINK_IDENTITY = 0, CGG_LINK_LOGIT = 1, CGG_LINK_LOG = 2, CGG_LINK_PROBIT = 3 };
/* beta_prior: distributional::dist_normal(mu, sigma) / dist_laplace(mu, sigma) / dist_student_t(df, mu, sigma) /
 * dist_gamma(shape, rate) / dist_exponential(rate) (R/glm_utils.R:108-110).  For gamma prior_mu = shape, prior_sigma = rate;
 * for exponential prior_sigma = rate.  A LIST of priors (R/glm_utils.R:113-115) is built with cgg_add_prior. */
enum { CGG_PRIOR_NORMAL = 0, CGG_PRIOR_LAPLACE = 1, CGG_PRIOR_STUDENT_T = 2, CGG_PRIOR_GAMMA = 3, CGG_PRIOR_EXPONENTIAL = 4 };
#define CGG_MAX_PRIORS 8
/* how the sweep is driven on the device */
enum {
    CGG_DRIVER_PERSISTENT = 0, /* one cooperative kernel runs all sweeps; per-chain flags, no grid barrier */
    CGG_DRIVER_STEPWISE = 1,   /* one launch per pass; the last CTA to finish decides */
    CGG_DRIVER_CLUSTER = 2     /* small n: one thread-block cluster per chain, sums exchanged through distributed shared
                                  memory, every CTA runs the chain's decisions itself (no global synchronisation).
                                  CGG_DRIVER_PERSISTENT selects it by itself for n <= 2^18 rows (CGG_SMALLN=0 disables) */
};
/* how rows are distributed */
enum {
    CGG_MODE_CHAINS = 0,      /* this handle holds all n rows; chains are independent (no collective) */
    CGG_MODE_ROW_SHARDED = 1  /* this handle holds one row shard; per-pass sums are exchanged via the
                                 exchange callback installed with cgg_set_exchange() */
};

/* cfg.flags */
#define CGG_FLAG_NO_PREFILTER 1 /* score every candidate in fp64 (the fp32 pre-filter never changes results, only cost) */
#define CGG_FLAG_NO_JET 2       /* decide every candidate from an exact pass over the rows (jet passes never change
                                   results, only cost: one pass per update instead of one per few candidates) */
#define CGG_FLAG_NAIVE 16       /* linear_predictor_calc = "naive" (R/glm_utils.R:206-208): eta is recomputed as X %*% beta, O(n p), before every
                                   pass over the rows instead of being updated in O(n); stepwise driver.  Same chain up to the rounding of eta */
#define CGG_FLAG_NO_CLUSTER 8   /* CGG_DRIVER_PERSISTENT: always the grid-wide kernel, also for small n (tests; never changes results) */
#define CGG_FLAG_NO_JET_LIGHT 4 /* binomial: every jet pass also evaluates the exact f(x0) (light passes skip it because
                                   the slice tests only involve differences f(v) - f(x0); never changes results) */

typedef struct cgg_config {
    int32_t abi_version; /* must be CGG_ABI_VERSION */
    int32_t device;      /* CUDA device ordinal */
    int64_t n;           /* rows held by this handle */
    int64_t p;           /* columns of the model matrix (intercept included) */
    int32_t family;      /* CGG_GAUSSIAN | CGG_BINOMIAL | CGG_POISSON */
    int32_t link;        /* must be the family's canonical link */
    double sd;           /* log_likelihood_extra_args$sd (R/mcmcglm.R:151); gaussian only */
    int32_t prior;       /* CGG_PRIOR_* */
    int32_t n_chains;    /* independent chains resident on this device (>= 1) */
    double prior_mu, prior_sigma, prior_df;
    double w;            /* qslice::slice_stepping_out `w` (> 0), R/mcmcglm.R:258-261 */
    int64_t max_steps;   /* qslice `max`; < 0 means Inf (the default) */
    int32_t K;           /* speculative candidates per chain per pass, 1..CGG_KMAX */
    int32_t driver;      /* CGG_DRIVER_* */
    int32_t mode;        /* CGG_MODE_* */
    int32_t chain_offset;/* global index of this handle's chain 0 (selects the Philox substream) */
    uint64_t seed;       /* Philox key */
    double spec_tau;     /* speculate a candidate only if P(needed) >= spec_tau; <= 0 => always fill K */
    int32_t rows_per_cta_min; /* 0 => default */
    int32_t flags;       /* CGG_FLAG_* */
    double jet_bound_scale; /* test knob: multiplies the jet enclosure's error bound (<= 0 => 1); any value >= 1 gives
                               the same chain, large values force the exact-pass fallback */
} cgg_config;

typedef struct cgg_stats {
    uint64_t updates;        /* coordinate updates completed (all chains) */
    uint64_t passes;         /* (chain, pass) pairs, idle ones included */
    uint64_t chain_passes;   /* (chain, pass) pairs that scored >= 1 candidate */
    uint64_t commit_passes;  /* (chain, pass) pairs that applied a pending eta update */
    uint64_t cand_evals;     /* candidates scored (speculative ones included) */
    uint64_t ref_evals;      /* evaluations qslice would have made (its nEvaluations), f(x0) included */
    uint64_t stepouts;       /* bracket expansions */
    uint64_t shrinks;        /* shrink proposals consumed (accepted one included) */
    uint64_t launches;       /* kernels launched by the last cgg_run */
    double sweep_ms;         /* device time of the last cgg_run's sweep kernels (CUDA events) */
    double algorithmic_bytes;/* 8n * (3*chain_passes + 2*commit_passes) of the last cgg_run */
    uint64_t coarse_evals;   /* candidates scored by the fp32 pre-filter (subset of cand_evals) */
    uint64_t coarse_undecided; /* passes that ended on a pre-filtered candidate the error bound could not decide */
    uint64_t jet_passes;     /* (chain, pass) pairs that were jet passes (subset of chain_passes) */
    uint64_t jet_fallbacks;  /* updates a jet pass could not finish: exact passes took over from that point */
    uint64_t jet_retries;    /* light jet passes (binomial) that were repeated as full jet passes */
    uint64_t group_passes;   /* walks over the rows that served FOUR chains at once, counted by one worker warp (cgg_run of the persistent
                              * driver; always 0 unless the library was built with -DCGG_GROUP_PASSES) */
} cgg_stats;

typedef struct cgg_handle cgg_handle;

/* Row-sharded exchange hook: called on the host between the local pass and the decision with a
 * DEVICE buffer of `count` doubles on `cuda_stream`; must sum it element-wise over all ranks in
 * place with a result that is bit-identical on every rank (e.g. all-gather + fixed-order sum). */
typedef int (*cgg_exchange_fn)(void *user, double *device_buf, int64_t count, void *cuda_stream);

const char *cgg_last_error(void);
int cgg_abi_version(void);

/* Replaces the front half of mcmcglm() that builds state (R/mcmcglm.R:171-198). */
int cgg_create(const cgg_config *cfg, cgg_handle **out);
void cgg_destroy(cgg_handle *h);

/* A list of priors (log_prior_density.list, R/glm_utils.R:113-115): the reference evaluates EVERY prior of the list at
 * EVERY coordinate and sums everything (quirk Q6), i.e. the coordinates are iid with density prod_k prior_k.  cfg.prior is
 * the first component; each call appends one (up to CGG_MAX_PRIORS in total): (a, b, c) = (mu, sigma, df) for normal /
 * laplace / student-t, (shape, rate, -) for gamma, (-, rate, -) for exponential.  Call before cgg_init_chain. */
int cgg_add_prior(cgg_handle *h, int32_t kind, double a, double b, double c);

/* X = model.matrix, Y = model.response (R/mcmcglm.R:176-178).  Host buffers, copied to HBM. */
int cgg_set_data(cgg_handle *h, const double *X_host, int64_t ldx, const double *y_host);
/* Same, but X/y already live in device memory (adopted, not copied; caller keeps them alive;
 * both 16-byte aligned, ldx even). */
int cgg_set_data_device(cgg_handle *h, const double *X_dev, int64_t ldx, const double *y_dev);

/* init_beta -> init_eta = X %*% init_beta (R/mcmcglm.R:200-216).  beta0 is drawn by the host
 * (the R shim uses distributional::generate exactly as the reference does). */
int cgg_init_chain(cgg_handle *h, int32_t chain, const double *beta0_host);

/* Sets a chain's state explicitly: current_beta[p] and current_eta[n] as the caller holds them (the
 * operator forms of R/glm_utils.R:126,187 take both as arguments; also the resume path).  eta is NOT
 * recomputed from beta. */
int cgg_set_state(cgg_handle *h, int32_t chain, const double *beta_host, const double *eta_host);

/* log_potential_from_betaj(new_beta_j, j, current_beta, current_eta, Y, X, family, beta_prior,
 * "update") of R/glm_utils.R:187-218 at K values of new_beta_j, against the chain's current
 * beta/eta.  Parity gate 1. */
int cgg_log_potential(cgg_handle *h, int32_t chain, int64_t j, int32_t K, const double *cand_host,
                      double *out_host);

/* update_linear_predictor(new_beta_j, current_beta_j, current_eta, X_j) of R/glm_utils.R:126-132
 * plus the commit of R/mcmcglm.R:264-268: beta[j] <- new_beta_j; eta <- eta + X_j * diff. */
int cgg_update_eta(cgg_handle *h, int32_t chain, int64_t j, double new_beta_j);

/* The (k, j) loop of R/mcmcglm.R:226-274 for n_iter further iterations of every chain.
 *   replay_u  NULL => on-device Philox; else n_chains streams of n_u uniforms each: replay_u[c*n_u + i] is the i-th
 *             uniform chain c consumes IN THIS CALL (every call brings its own buffer), consumed exactly as qslice's
 *             runif(1) calls would.  Running out of them is CGG_E_STREAM.
 *   u_consumed  [n_chains] cumulative uniforms consumed per chain since it was initialised (nullable).
 *   samples_out [n_chains][n_iter][p] row-major: beta after each iteration (nullable).
 * May be called repeatedly; the chain state persists between calls.  The Philox stream of a chain is indexed by that
 * cumulative count, which cgg_init_chain / cgg_set_state reset to 0: re-initialising a chain replays the same numbers
 * unless cfg.seed or cfg.chain_offset differ.
 * After a device-side failure of a chain (CGG_E_NAN, CGG_E_STREAM, CGG_E_NOTERM) that chain stopped in the middle of an
 * update; it is marked uninitialised and must be given a state again (cgg_init_chain / cgg_set_state) before the next
 * cgg_run, which otherwise returns CGG_E_STATE. */
int cgg_run(cgg_handle *h, int64_t n_iter, const double *replay_u, uint64_t n_u, uint64_t *u_consumed,
            double *samples_out, cgg_stats *stats);

/* qslice's tuning parameter per chain: w_host[n_chains] replaces cfg.w from the next cgg_run on.  This is how
 * mcmcglm_across_tuningparams (R/slice_utilities.R:43-85) becomes ONE engine run: the chains share every X_j read. */
int cgg_set_chain_w(cgg_handle *h, const double *w_host);
/* The counters of one chain from the last cgg_run (ref_evals = qslice's nEvaluations, which the reference drops at
 * R/mcmcglm.R:261); launches and sweep_ms are not per chain and stay 0. */
int cgg_get_chain_stats(cgg_handle *h, int32_t chain, cgg_stats *out);

/* Current beta[p] and (nullable) eta[n] of a chain, to host. */
int cgg_get_state(cgg_handle *h, int32_t chain, double *beta_host, double *eta_host);
/* Carried log-potential at the chain's current point (what qslice's first f(x) would return). */
int cgg_get_fx(cgg_handle *h, int32_t chain, double *fx);

/* Diagnostic: out[i] = the per-row log-density term the kernels accumulate for (y_i, eta_i), i.e.
 * log_density(family, linkinv(eta_i), y_i) of R/glm_utils.R:24-57 minus the terms that do not depend on
 * beta (gaussian: -log(sqrt(2 pi) sd); poisson: -lgamma(y + 1)).  Used by the accuracy tests. */
int cgg_debug_row_terms(int32_t device, int32_t family, int64_t n, const double *y_host, const double *eta_host,
                        double sd, double *out_host);

/* Diagnostic: exhaustive scan over every fp32 s with |s| <= 37 of the fp32 pre-filter's softplus against the
 * fp64 one; returns max |err| / (1 + |s|), the constant its error bound relies on (pre-filter). */
int cgg_debug_coarse_error(int32_t device, double *max_err_over_1_plus_abs_s, double *at_s);

/* Diagnostic: max absolute error of the per-row quantities of a binomial LIGHT jet pass (tanh(|eta|/2), s(1-s) and their
 * product, evaluated with a 1e-14 exp and a one-step reciprocal) against double-precision library routines over a grid of
 * 6.7e7 values of |eta| in [0, 40); the enclosure's error bound budgets 2e-11 for it. */
int cgg_debug_light_error(int32_t device, double *max_abs_err);

/* Diagnostic: runs one jet pass of chain `chain` along column j (no state change) and evaluates its enclosure at K
 * (<= CGG_KMAX) values of new_beta_j: value[k] = surrogate log-LIKELIHOOD (per-dataset constant included, prior
 * excluded), bound[k] = the error bound the decider uses (Inf: enclosure not applicable), sums[CGG_KMAX + 2] = the
 * pass's raw sums (nullable).  light != 0 (binomial only): a light pass; value[k] is then the log-likelihood
 * DIFFERENCE to the current point.  Used by the tests that check the bound against exact evaluations. */
int cgg_debug_jet(cgg_handle *h, int32_t chain, int64_t j, int32_t K, int32_t light, const double *cand_host,
                  double *value_host, double *bound_host, double *sums_host);

int cgg_set_exchange(cgg_handle *h, cgg_exchange_fn fn, void *user);

/* Built-in exchange for row-sharded handles over NCCL (NVLink / NVSwitch): all-gather of the per-candidate
 * partial sums followed by a sum in rank order, so every rank sees bit-identical totals.  NCCL is resolved
 * at run time (dlopen "libnccl.so.2"); rank 0 creates the id with cgg_nccl_unique_id and the host code
 * distributes the 128 bytes to the other ranks by whatever means it has. */
#define CGG_NCCL_ID_BYTES 128
int cgg_nccl_unique_id(char out[CGG_NCCL_ID_BYTES]);
int cgg_comm_init_nccl(cgg_handle *h, int32_t rank, int32_t world, const char id[CGG_NCCL_ID_BYTES]);
/* Row-sharded handles on CGG_DRIVER_PERSISTENT: the exchange of a pass's sums stays INSIDE the persistent kernel.  Every
 * rank owns a mailbox in its device memory; after a pass the deciding warp of each rank stores the rank's sums into every
 * peer's mailbox (peer memory: NVLink / NVSwitch, system-scope stores of self-validating words), polls its own mailbox and
 * adds the ranks' parts in rank order: bit-identical totals, hence identical decisions, on every rank; no collective call,
 * no host round trip, no extra launch per pass.
 *   cgg_p2p_mailbox  allocates this handle's mailbox for `world` ranks (<= 8); returns its device pointer and / or its CUDA
 *                    IPC handle (64 bytes), which the host code hands to the other ranks by whatever means it has.
 *   cgg_p2p_connect  maps the peers' mailboxes: dev_ptrs[r] (nullable array; a pointer valid in THIS process, e.g. another
 *                    handle on the same device) or else ipc_handles + 64 r (another process).  Entry `rank` is ignored.
 * Column statistics of the jet passes are still reduced once at cgg_set_data through cgg_comm_init_nccl / cgg_set_exchange. */
#define CGG_IPC_HANDLE_BYTES 64
int cgg_p2p_mailbox(cgg_handle *h, int32_t world, void **dev_ptr, char ipc_handle[CGG_IPC_HANDLE_BYTES]);
int cgg_p2p_connect(cgg_handle *h, int32_t rank, int32_t world, void *const *dev_ptrs, const char *ipc_handles);

/* CUDA stream (cudaStream_t) the handle launches on, for callers that time with their own events */
void *cgg_stream(cgg_handle *h);
/* Grid used by the sweep kernels: CTAs and threads per CTA (for launch accounting) */
int cgg_launch_shape(cgg_handle *h, int32_t *ctas, int32_t *threads);

#ifdef __cplusplus
}
#endif
#endif /* CGGIBBS_H */
