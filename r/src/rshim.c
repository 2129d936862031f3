/* rshim.c -- .Call layer between R and libcggibbs (include/cggibbs.h).  One function per C-ABI entry
 * point, nothing but SEXP unwrapping: all logic lives behind the C ABI where the ctypes tests reach it.
 * EXPERIMENTAL: the image this repository is built in has no R, so this file has only ever been syntax-checked
 * (tests/test_abi_cpu.py compiles it with -fsyntax-only against the stub headers in tests/r_stub); see INTEGRATION.md. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "cggibbs.h"

static void chk(int rc) { if (rc != CGG_OK) Rf_error("%s", cgg_last_error()); }

static void handle_finalizer(SEXP p) {
    cgg_handle *h = (cgg_handle *)R_ExternalPtrAddr(p);
    if (h) { cgg_destroy(h); R_ClearExternalPtr(p); }
}
static cgg_handle *get_handle(SEXP p) {
    cgg_handle *h = (cgg_handle *)R_ExternalPtrAddr(p);
    if (!h) Rf_error("cggibbs handle has been destroyed");
    return h;
}
static double num(SEXP lst, const char *name, double dflt) {
    SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(lst, i));
    return dflt;
}

/* cfg: named list built by R/engine.R:cgg_config() */
SEXP C_cgg_create(SEXP cfg) {
    cgg_config c;
    memset(&c, 0, sizeof c);
    c.abi_version = CGG_ABI_VERSION;
    c.device = (int32_t)num(cfg, "device", 0);
    c.n = (int64_t)num(cfg, "n", 0); c.p = (int64_t)num(cfg, "p", 0);
    c.family = (int32_t)num(cfg, "family", -1); c.link = (int32_t)num(cfg, "link", -1);
    c.sd = num(cfg, "sd", 1.0);
    c.prior = (int32_t)num(cfg, "prior", -1);
    c.prior_mu = num(cfg, "prior_mu", 0.0); c.prior_sigma = num(cfg, "prior_sigma", 1.0); c.prior_df = num(cfg, "prior_df", 1.0);
    c.n_chains = (int32_t)num(cfg, "n_chains", 1);
    c.w = num(cfg, "w", NA_REAL);
    c.max_steps = (int64_t)num(cfg, "max_steps", -1);
    c.K = (int32_t)num(cfg, "K", 8);
    c.driver = (int32_t)num(cfg, "driver", CGG_DRIVER_PERSISTENT);
    c.mode = CGG_MODE_CHAINS;
    c.chain_offset = (int32_t)num(cfg, "chain_offset", 0);
    c.seed = (uint64_t)num(cfg, "seed", 0);
    c.spec_tau = num(cfg, "spec_tau", 0.12);
    c.flags = (int32_t)num(cfg, "flags", 0);
    c.jet_bound_scale = 0.0;
    cgg_handle *h = NULL;
    chk(cgg_create(&c, &h));
    SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, handle_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* one more prior of a list of priors (R/glm_utils.R:113-115): kind and its (a, b, c) */
SEXP C_cgg_add_prior(SEXP p, SEXP kind, SEXP a, SEXP b, SEXP c) {
    chk(cgg_add_prior(get_handle(p), Rf_asInteger(kind), Rf_asReal(a), Rf_asReal(b), Rf_asReal(c)));
    return R_NilValue;
}
/* per-chain slice widths (mcmcglm_across_tuningparams as one engine run) */
SEXP C_cgg_set_chain_w(SEXP p, SEXP w) {
    chk(cgg_set_chain_w(get_handle(p), REAL(w)));
    return R_NilValue;
}
/* X: the double model matrix (already column-major with ld = nrow), y: double vector */
SEXP C_cgg_set_data(SEXP p, SEXP X, SEXP y) {
    chk(cgg_set_data(get_handle(p), REAL(X), (int64_t)Rf_nrows(X), REAL(y)));
    return R_NilValue;
}
SEXP C_cgg_init_chain(SEXP p, SEXP chain, SEXP beta0) {
    chk(cgg_init_chain(get_handle(p), Rf_asInteger(chain) - 1, REAL(beta0)));
    return R_NilValue;
}
SEXP C_cgg_set_state(SEXP p, SEXP chain, SEXP beta, SEXP eta) {
    chk(cgg_set_state(get_handle(p), Rf_asInteger(chain) - 1, REAL(beta), REAL(eta)));
    return R_NilValue;
}
/* j is 1-based on the R side */
SEXP C_cgg_log_potential(SEXP p, SEXP chain, SEXP j, SEXP cand) {
    R_xlen_t K = XLENGTH(cand);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, K));
    chk(cgg_log_potential(get_handle(p), Rf_asInteger(chain) - 1, (int64_t)Rf_asInteger(j) - 1, (int32_t)K, REAL(cand), REAL(out)));
    UNPROTECT(1);
    return out;
}
SEXP C_cgg_update_eta(SEXP p, SEXP chain, SEXP j, SEXP new_beta_j) {
    chk(cgg_update_eta(get_handle(p), Rf_asInteger(chain) - 1, (int64_t)Rf_asInteger(j) - 1, Rf_asReal(new_beta_j)));
    return R_NilValue;
}
SEXP C_cgg_get_state(SEXP p, SEXP chain, SEXP n, SEXP np) {
    SEXP beta = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(np)));
    SEXP eta = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(n)));
    chk(cgg_get_state(get_handle(p), Rf_asInteger(chain) - 1, REAL(beta), REAL(eta)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, beta); SET_VECTOR_ELT(out, 1, eta);
    UNPROTECT(3);
    return out;
}
/* replay_u: NULL, or a double matrix with one COLUMN per chain: the runif draws of THIS call (the front end calls in
 * chunks of a few iterations, so that the progress bar moves and Ctrl-C is honoured between chunks) */
SEXP C_cgg_run(SEXP p, SEXP n_iter, SEXP n_chains, SEXP np, SEXP replay_u) {
    R_CheckUserInterrupt();
    int C = Rf_asInteger(n_chains), P = Rf_asInteger(np);
    int64_t it = (int64_t)Rf_asReal(n_iter);
    SEXP smp = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C * it * P));   /* [chain][iteration][coef], row-major */
    SEXP used = PROTECT(Rf_allocVector(REALSXP, C));
    uint64_t *u = (uint64_t *)R_alloc(C, sizeof(uint64_t));
    cgg_stats st;
    const double *ru = Rf_isNull(replay_u) ? NULL : REAL(replay_u);
    uint64_t n_u = Rf_isNull(replay_u) ? 0 : (uint64_t)Rf_nrows(replay_u);
    chk(cgg_run(get_handle(p), it, ru, n_u, u, REAL(smp), &st));
    for (int c = 0; c < C; ++c) REAL(used)[c] = (double)u[c];
    const char *nm[] = {"samples", "uniforms_used", "updates", "passes", "cand_evals", "ref_evals", "stepouts", "shrinks", "sweep_ms", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, smp); SET_VECTOR_ELT(out, 1, used);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)st.updates)); SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)st.passes));
    SET_VECTOR_ELT(out, 4, Rf_ScalarReal((double)st.cand_evals)); SET_VECTOR_ELT(out, 5, Rf_ScalarReal((double)st.ref_evals));
    SET_VECTOR_ELT(out, 6, Rf_ScalarReal((double)st.stepouts)); SET_VECTOR_ELT(out, 7, Rf_ScalarReal((double)st.shrinks));
    SET_VECTOR_ELT(out, 8, Rf_ScalarReal(st.sweep_ms));
    UNPROTECT(3);
    return out;
}
SEXP C_cgg_destroy(SEXP p) { handle_finalizer(p); return R_NilValue; }

static const R_CallMethodDef call_methods[] = {
    {"C_cgg_create", (DL_FUNC)&C_cgg_create, 1},          {"C_cgg_set_data", (DL_FUNC)&C_cgg_set_data, 3},
    {"C_cgg_init_chain", (DL_FUNC)&C_cgg_init_chain, 3},  {"C_cgg_set_state", (DL_FUNC)&C_cgg_set_state, 4},
    {"C_cgg_log_potential", (DL_FUNC)&C_cgg_log_potential, 4}, {"C_cgg_update_eta", (DL_FUNC)&C_cgg_update_eta, 4},
    {"C_cgg_get_state", (DL_FUNC)&C_cgg_get_state, 4},    {"C_cgg_run", (DL_FUNC)&C_cgg_run, 5},
    {"C_cgg_destroy", (DL_FUNC)&C_cgg_destroy, 1},        {"C_cgg_add_prior", (DL_FUNC)&C_cgg_add_prior, 5},
    {"C_cgg_set_chain_w", (DL_FUNC)&C_cgg_set_chain_w, 2}, {NULL, NULL, 0}};

void R_init_mcmcglmb200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
