// cgg_device.cuh -- device-side data layout and the device routines every sweep kernel is made of:
//   warp_pass_jet<FAMILY, FULL>()  : one warp, one chain, one JET pass over the warp's row tiles: applies the pending eta
//                                    update and accumulates the derivative moments of the log-likelihood along the
//                                    column (cgg_jet.cuh); warp_pass_jet2: the same for two chains at the same coordinate
//   warp_pass_chain<FAMILY>()      : one warp, one chain, one EXACT pass: scores every live candidate in fp64 (fp32
//                                    pre-filter for far ones) and applies a pending eta update (reference a4..a9 + a5)
//   jet_decide()                   : the whole qslice stepping-out / shrinkage update (reference a3) from one jet pass
//   decide_chain()                 : one decision of a chain: jet_decide, or the resumable state machine over exact passes
// Workers are WARPS.  Worker w owns the 64-row tiles w, w+W, w+2W, ... for the whole run (at any moment the grid touches
// one contiguous window of every operand); chains are independent pipelines pass -> decision -> pass with their own
// result slots and version flag, so there is no grid-wide barrier anywhere.
#pragma once
#include "cgg_math.cuh"
#include "cgg_jet.cuh"

namespace cgg {

constexpr int KMAX = CGG_KMAX;
constexpr int NV = KMAX + 2;    // values a chain pass delivers: KMAX candidate sums + the two error-bound sums of the pre-filter
#ifndef CGG_THREADS
#define CGG_THREADS 256     // 8 warps per SM: measured best for the round-2 kernels (cfg3: 128 -> 186k, 192 -> 199k, 256 -> 222k, 320 -> 203k,
                            // 384 -> 203k, 512 -> 154k updates/s): every warp pays the per-pass fixed work (reductions, delivery, control)
#endif
#ifndef CGG_RING_D
#define CGG_RING_D 4
#endif
#if !defined(CGG_LEAN_PAIR) && !defined(CGG_NO_LEAN_PAIR)
#define CGG_LEAN_PAIR 1      // pair passes of the steady state run through the branch-free GroupStream<2> loop (cggibbs.cu, `kind == 3`):
                             // 204 instead of 238 SASS instructions per tile, +12 % on the headline workload; -DCGG_NO_LEAN_PAIR: off
#endif
#ifndef CGG_PAIR_TPI
#define CGG_PAIR_TPI 1       // tiles per iteration of the pair-pass loop (1 or 2)
#endif
#ifndef CGG_JET_TPI
#define CGG_JET_TPI 2        // tiles a warp scores per iteration of the jet loop (1 or 2)
#endif
constexpr int THREADS = CGG_THREADS;   // one CTA per SM, THREADS/32 warp-workers each
constexpr int NWARPS = THREADS / 32;
constexpr int CMAX = 32;       // chains per device (shared-memory slots of the CTA-level reduction)
constexpr int NU = 12;         // uniforms fetched per decision: 3 start draws + KMAX proposals (+1 spare)
constexpr int RING_D = CGG_RING_D;  // tiles in flight per warp (cp.async ring depth)
constexpr int RING_OPS = 5;    // eta, y, X_j, X_commit (single pass) | eta_A, eta_B, y, X_j, X_commit (pair pass)
constexpr int TILE_ROWS = 64;  // 32 lanes x one 128-bit transfer per operand
constexpr int RING_BYTES_PER_WARP = RING_D * RING_OPS * 32 * 16;

enum Phase : int32_t { PH_START = 0, PH_STEPOUT = 1, PH_SHRINK = 2, PH_FLUSH = 3, PH_FINISHED = 4, PH_JET = 5 };
static_assert(JET_NV == NV, "a jet pass delivers through the same accumulators as a candidate pass");
constexpr unsigned JET_BIT = 0x80000000u;   // Ctl::coarse_mask: this pass is a jet pass (cgg_jet.cuh), ncand = 0
constexpr unsigned JET_FULL = 0x40000000u;  // ... that also delivers the exact M_0 (always, except binomial light passes)

// What every worker needs to know about a chain for its coming pass.  12 x 8 bytes.
struct __align__(16) Ctl {
    int32_t j;          // column being sampled; -1: chain finished or failed, skip it for good
    int32_t ncand;      // candidates to score (0..KMAX)
    int32_t commit_j;   // column of a pending eta update to apply first (-1: none)
    int32_t coarse_mask;// bit k: candidate k is scored by the fp32 pre-filter (binomial only)
    double commit_delta;// new_beta_j - current_beta_j of that update (R/glm_utils.R:127)
    double delta[KMAX]; // cand_k - beta_j
    double cscale;      // jet pass: the column's power-of-two scale (colstat[j][0]), so workers need no dependent load
};
static_assert(sizeof(Ctl) == 96, "Ctl must be 12 doubles");
constexpr int CTL_WORDS = sizeof(Ctl) / 8;

// Per-chain synchronisation words, one 128-byte line each.
struct __align__(128) ChainSync {
    unsigned long long arrive;   // monotone: CTAs that finished a pass of this chain
    unsigned long long version;  // passes decided so far (release/acquire flag)
    unsigned long long pad[14];
};

// Exact accumulator of one candidate's log-likelihood: 128-bit two's-complement fixed point with 64
// fractional bits, updated with integer atomics.  Integer addition is associative, so the total does not
// depend on arrival order: sums are bit-reproducible without a serial reduction over per-CTA partials.
struct __align__(128) Acc {  // one L2 line each, so different candidates hit different L2 slices
    unsigned long long lo;
    long long hi;
    unsigned int flags;  // 1: a partial was -Inf, 2: NaN, 4: +Inf or out of range
    unsigned int pad[27];
};

// Slice-sampler state of one chain; touched only by the deciding warp.
struct __align__(16) ChainState {
    double x0, fx0, ylev, L, R, Jb, Kb;
    double prior_sum;   // sum_l log pi(beta_l) at the current beta
    double prior_rest;  // prior_sum - log pi(beta_j)
    double pexp;        // running estimate of P(a stepping-out expansion happens)
    double cand[KMAX];
    int32_t phase, status;
    int32_t nL, nR, nS;
    int32_t openL, openR;
    int32_t sdrawn;     // shrink uniforms already consumed by rejected proposals of this update
    int32_t npass;      // passes spent on this update (non-termination guard)
    int32_t j;
    int32_t fine_next;  // the previous pass left a pre-filtered candidate undecided: score everything in fp64 next
    int32_t jet_skip;   // a jet pass of this sweep met rows within reach of a clamp: exact passes until the sweep ends
    int64_t iter;       // iterations completed in this run
    uint64_t cursor;    // uniforms consumed before this update
    uint64_t updates, chain_passes, commit_passes, cand_evals, ref_evals, stepouts, shrinks, passes;
    uint64_t coarse_evals, coarse_undecided;
    uint64_t jet_passes, jet_fallbacks, jet_retries;
    double w;           // this chain's slice width (qslice's `w`; per chain so that a tuning sweep is ONE engine run)
};
static_assert(sizeof(ChainState) % 16 == 0, "ChainState is copied with 128-bit accesses");

struct Hdr {
    unsigned int ticket;        // stepwise driver: CTAs finished in this launch
    int32_t done;               // stepwise driver: every chain finished (or failed)
    int32_t abort;              // a wait timed out
    int32_t group_passes;       // persistent driver: group passes (four chains per walk) walked by worker warp 0 in this cgg_run
};

struct LimbAcc;
struct Dev {
    const double *X; const double *y;
    double *eta, *beta, *shat, *samples, *xbuf;
    const double *replay;
    const double *colstat;     // [p][CS_STRIDE] per-column scale and absolute moments (jet passes)
    const double *gathered;    // row-sharded over NCCL: [world][C * NV] all-gathered per-rank sums (added in rank order by the decider)
    LimbAcc *lacc;             // [C][NV] limb accumulators of the pass in flight (persistent driver), see cta_deliver_limbs
    Ctl *ctl; ChainState *cs; Hdr *hdr; ChainSync *sync; Acc *acc;
    unsigned long long *prof;  // optional phase counters (CGG_PROFILE=1)
    int64_t n, p, ldx, lde, n_tiles, n_iter, iter_stop;   // iter_stop: iteration count at which this launch stops (<= n_iter)
    uint64_t n_u, seed;
    int64_t max_steps;
    double inv_sd, ll_const, w, tau, coarse_theta, jet_bscale, n_total;   // n_total: rows of all shards (error bounds); w: default slice width (ChainState::w is what the chains use)
    double jet_ce;             // rounding allowance of one accumulated moment, relative to the sum of its terms' magnitudes (cgg_create: from the summation depth)
    // row-sharded persistent driver: peer mailboxes (one per rank, each [C][world][NV] stamped 16-byte entries); mbox[r] is rank r's
    // mailbox as mapped into THIS process (peer memory over NVLink, or the same device), nullptr: the exchange is not through mailboxes
    unsigned long long *mbox[8];
    uint32_t mbox_stamp0;      // stamps already used by earlier runs of this handle (all ranks agree: they run identical passes)
    int32_t rank;
    uint64_t replay_origin[CMAX];   // replay mode: the chain's uniform cursor at the start of this cgg_run (replay_u is indexed from there)
    PriorSet prior;
    int32_t C, K, G, family, chain_offset, sharded, coarse, jet, jet_light, world, pair, colcache, quad, early;   // pair: chains 2k, 2k+1 share a pass when they can; quad: chains 4k .. 4k+3 do (needs pair and the cache);
                                                                                                  // colcache: tiles per slot of the per-warp X-column cache (0: off)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// acquire-release fence at device scope (cheaper than __threadfence(), which is a sequentially consistent MEMBAR.SC)
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- asynchronous global -> shared staging (LDGSTS), L2-coherent (.cg bypasses L1) ---------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// the same under a predicate, in ONE instruction stream: no branch around the copy, and consecutive predicated copies stay
// one LDGSTS group for ptxas (which puts three filler instructions in front of every group)
__device__ __forceinline__ void cp_async16_if(bool p, uint32_t smem_addr, const void *gptr) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(smem_addr), "l"(gptr), "r"((int)p) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double2 lds2(uint32_t smem_addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_addr) : "memory");
    return v;
}

// ---- exact accumulation ---------------------------------------------------------------------
__device__ __forceinline__ void acc_add(Acc *a, double v) {
    if (!(fabs(v) < 9.0e18)) {  // -Inf, NaN, +Inf or beyond the fixed-point range
        atomicOr(&a->flags, (v != v) ? 2u : ((v < 0.0) ? 1u : 4u));
        return;
    }
    const double fl = floor(v);
    const long long hi = (long long)fl;
    const unsigned long long lo = (unsigned long long)((v - fl) * 18446744073709551616.0);  // (v - fl) in [0,1)
    long long carry = 0;
    if (lo) {
        const unsigned long long old = atomicAdd(&a->lo, lo);
        carry = (old + lo < old) ? 1 : 0;
    }
    if (hi + carry) atomicAdd(reinterpret_cast<unsigned long long *>(&a->hi), (unsigned long long)(hi + carry));
}

// Read-and-clear by the single deciding lane (all adds of this pass are ordered before by the arrive counter).
__device__ __forceinline__ double acc_take(Acc *a, unsigned int *flags_out = nullptr) {
    const unsigned long long lo = __ldcg(&a->lo);
    const long long hi = __ldcg(&a->hi);
    const unsigned int fl = __ldcg(&a->flags);
    if (flags_out) *flags_out = fl;
    a->lo = 0ULL; a->hi = 0LL; a->flags = 0u;
    if (fl & 2u) return NAN;
    if ((fl & 1u) && (fl & 4u)) return NAN;
    if (fl & 1u) return -INFINITY;
    if (fl & 4u) return INFINITY;
    return (double)hi + (double)lo * 5.421010862427522e-20;  // 2^-64
}

// ---------------------------------------------------------------------------------------------
// Streaming of one chain's operands by one warp.  The warp owns tiles T = vw, vw + W, ... (64 rows each, one
// 128-bit transfer per lane per operand).  Algorithmic traffic: 8 B/row each of y, eta, X_j; a pending commit
// adds X_commit (read) and eta (write).  Operand tiles are staged global -> shared with cp.async into a
// RING_D-deep per-warp ring, so RING_D - 1 tiles are in flight while one is being scored; every lane reads
// back only the 16-byte slots it copied itself, so no barrier of any kind is needed.
struct ChainStream {
    const double *xj, *xc, *y;
    double *eta;
    long long vw, W, n_tiles;
    int64_t n;
    int nc, cj;
    bool need_yx;            // the pass reads y and X_j (candidates to score, or a jet pass)
    uint32_t slot0;
    // The tile -> worker map is rotated per chain (fixed for the run): n_tiles is rarely a multiple of W, so some
    // workers own one tile more than others; rotating by c * W / C spreads those extra tiles evenly over the
    // workers within a round of C chains, which is the granularity at which workers have slack.
    // OWNERSHIP INVARIANT (the protocol's only ordering of eta accesses): for the whole run, row i of chain c's eta is
    // read and written by one and the same warp -- `vw` below is the only tile -> warp map and PairStream computes the
    // same value -- so eta never needs a fence or a release between passes.
    __device__ __forceinline__ ChainStream(const Dev &d, int c, const double *cw, long long wid, long long W_, int lane, uint32_t ring) {
        const long long w0 = __double_as_longlong(cw[0]), w1 = __double_as_longlong(cw[1]);
        const int j = (int)(w0 & 0xffffffffLL);
        nc = (int)(w0 >> 32); cj = (int)(w1 & 0xffffffffLL);
        need_yx = nc > 0 || (((unsigned)(w1 >> 32)) & JET_BIT);
        xj = d.X + (int64_t)(j < 0 ? 0 : j) * d.ldx;
        xc = d.X + (int64_t)(cj < 0 ? 0 : cj) * d.ldx;
        y = d.y; eta = d.eta + (int64_t)c * d.lde;
        // with pair passes on, both chains of a pair use the even chain's rotation in EVERY kind of pass (PairStream does the
        // same), so a chain's eta rows keep their owner warp when it moves between pair passes and single passes
        // ... and with the X-column cache on (ColCache) there is no rotation at all: a warp's cached tiles serve every pair
        W = W_; vw = d.colcache ? wid : (wid + (long long)(d.pair ? (c & ~1) : c) * (W / d.C)) % W; n_tiles = d.n_tiles; n = d.n;
        slot0 = ring + (uint32_t)lane * 16u;
    }
    __device__ __forceinline__ void issue(long long T, int stage, int lane) const {
        const int64_t i = T * TILE_ROWS + 2 * lane;
        if (T < n_tiles && i + 1 < n) {
            const uint32_t s = slot0 + (uint32_t)stage * (RING_OPS * 512u);
            cp_async16(s, eta + i);
            if (need_yx) { cp_async16(s + 512u, y + i); cp_async16(s + 1024u, xj + i); }
            if (cj >= 0) cp_async16(s + 1536u, xc + i);
        }
        cp_async_commit();
    }
    // same, addressed by the row index i = T * TILE_ROWS + 2 * lane of this lane's pair (i + 1 < n implies T < n_tiles)
    __device__ __forceinline__ void issue_at(int64_t i, int stage) const {
        if (i + 1 < n) {
            const uint32_t s = slot0 + (uint32_t)stage * (RING_OPS * 512u);
            cp_async16(s, eta + i);
            if (need_yx) { cp_async16(s + 512u, y + i); cp_async16(s + 1024u, xj + i); }
            if (cj >= 0) cp_async16(s + 1536u, xc + i);
        }
        cp_async_commit();
    }
    // first RING_D - 1 tiles; may be issued early, while the previous chain is still being reduced
    __device__ __forceinline__ void prologue(int lane) const {
#pragma unroll
        for (int s = 0; s < RING_D - 1; ++s) issue(vw + s * W, s, lane);
    }
};

// Scores every live candidate of the chain on NP row pairs (NP tiles of this lane) and adds the terms to the per-lane
// running sums in shared memory.  Two candidates per loop iteration x NP tiles x two rows: the independent dependency
// chains a warp needs in flight with 2 warps per scheduler.
template <int FAMILY, int NP>
__device__ __forceinline__ void score_pairs(const RowPair<FAMILY> (&rp)[NP], int nc, unsigned cmask, const double *s_dl, double inv_sd,
                                            const double2 *tab, double *sacc, int lane, float &bE, float &bX, unsigned &nearmask, float &smax) {
    const unsigned all = (nc >= 32) ? 0xffffffffu : ((1u << nc) - 1u);
    unsigned fine = all & ~cmask;
    if constexpr (FAMILY == CGG_BINOMIAL) if (cmask) {
        // pre-filter: candidates flagged in cmask are scored in fp32 (hardware ex2/lg2); the sums of
        // |eta| and |x| over the rows feed the rigorous error bound the decider applies
        float ef0[NP], ef1[NP], xf0[NP], xf1[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            ef0[q] = (float)rp[q].e0; ef1[q] = (float)rp[q].e1; xf0[q] = (float)rp[q].x0; xf1[q] = (float)rp[q].x1;
            bE += fabsf(ef0[q]) + fabsf(ef1[q]);
            bX += fabsf(xf0[q]) + fabsf(xf1[q]);
        }
        unsigned cm = cmask & all;
        while (cm) {
            const int k0 = __ffs(cm) - 1; cm &= cm - 1;
            const int k1 = cm ? __ffs(cm) - 1 : k0;
            if (cm) cm &= cm - 1;
            const float d0 = (float)s_dl[k0], d1 = (float)s_dl[k1];
            bool n0 = false, n1 = false;
            float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const float s00 = fmaf(xf0[q], d0, ef0[q]), s01 = fmaf(xf1[q], d0, ef1[q]);
                const float s10 = fmaf(xf0[q], d1, ef0[q]), s11 = fmaf(xf1[q], d1, ef1[q]);
                // smax doubles as the accumulator of the bound on R's log(1 - p) rounding: sum of e^s over the (row, candidate)
                // pairs in that regime (valid -- if generous -- for every pre-filtered candidate of the pass)
                v0 += softplus32n(s00, rp[q].z0, n0, smax) + softplus32n(s01, rp[q].z1, n0, smax);
                v1 += softplus32n(s10, rp[q].z0, n1, smax) + softplus32n(s11, rp[q].z1, n1, smax);
            }
            if (n0) nearmask |= 1u << k0;
            sacc[k0 * 32 + lane] -= (double)v0;
            if (k1 != k0) { if (n1) nearmask |= 1u << k1; sacc[k1 * 32 + lane] -= (double)v1; }
        }
    }
    while (fine) {
        const int k0 = __ffs(fine) - 1; fine &= fine - 1;
        if (fine) {
            const int k1 = __ffs(fine) - 1; fine &= fine - 1;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int q = 0; q < NP; ++q) { t0 += rp[q].term(s_dl[k0], inv_sd, tab); t1 += rp[q].term(s_dl[k1], inv_sd, tab); }
            sacc[k0 * 32 + lane] += t0;
            sacc[k1 * 32 + lane] += t1;
        } else {
            double t0 = 0.0;
#pragma unroll
            for (int q = 0; q < NP; ++q) t0 += rp[q].term(s_dl[k0], inv_sd, tab);
            sacc[k0 * 32 + lane] += t0;
        }
    }
}

// One warp, one chain, one pass.  The candidate loop is a run-time loop (trip count nc is warp-uniform) with
// the per-lane running sums in shared memory: sacc[k * 32 + lane].  s_dl[k] = cand_k - beta_j comes from the
// CTA's shared control block.  `prefetched`: the prologue of this chain was already issued.  Two tiles per iteration.
template <int FAMILY>
__device__ __forceinline__ void warp_pass_chain(const Dev &d, const ChainStream &cs, double cdelta, const double *s_dl,
                                                int lane, const double2 *tab, double *sacc, bool prefetched,
                                                unsigned cmask, float &bE, float &bX, unsigned &nearmask) {
    const int nc = cs.nc, cj = cs.cj;
    const int64_t n = cs.n;
    double *eta = cs.eta;
    float smax = 0.0f;          // pre-filter: sum of e^s over this lane's (row, candidate) pairs in R's log(1 - p) rounding regime
    unsigned nrows = 0;         // rows this lane scored
    if (!prefetched) cs.prologue(lane);
    for (int k = 0; k < nc; ++k) sacc[k * 32 + lane] = 0.0;
    static_assert(RING_D >= 4 && (RING_D & (RING_D - 1)) == 0, "two tiles per iteration: ring of >= 4 stages, power of two");
    // the committed (and, if an update is pending, written back) eta of this lane's pair in tile T
    auto load_eta = [&](unsigned stage, int64_t i) {
        const uint32_t s = cs.slot0 + stage * (RING_OPS * 512u);
        double2 e = lds2(s);
        if (cj >= 0) {
            const double2 cv = lds2(s + 1536u);
            e.x = eta_shift(e.x, cv.x, cdelta);
            e.y = eta_shift(e.y, cv.y, cdelta);
            *reinterpret_cast<double2 *>(eta + i) = e;
        }
        return e;
    };
    unsigned stage = 0;
    for (long long T = cs.vw; T < cs.n_tiles; T += 2 * cs.W) {
        cs.issue(T + (RING_D - 1) * cs.W, (int)((stage + RING_D - 1) & (RING_D - 1)), lane);
        cp_async_wait<RING_D - 2>();
        const unsigned st1 = (stage + 1) & (RING_D - 1);
        const int64_t i0 = T * TILE_ROWS + 2 * lane, i1 = i0 + cs.W * TILE_ROWS;
        if (i1 + 1 < n) {
            const double2 ea = load_eta(stage, i0), eb = load_eta(st1, i1);
            if (nc > 0) {
                const uint32_t sa = cs.slot0 + stage * (RING_OPS * 512u), sb = cs.slot0 + st1 * (RING_OPS * 512u);
                const RowPair<FAMILY> rp[2] = {RowPair<FAMILY>(lds2(sa + 512u), ea, lds2(sa + 1024u)),
                                               RowPair<FAMILY>(lds2(sb + 512u), eb, lds2(sb + 1024u))};
                score_pairs<FAMILY, 2>(rp, nc, cmask, s_dl, d.inv_sd, tab, sacc, lane, bE, bX, nearmask, smax);
                nrows += 4;
            }
        } else if (i0 + 1 < n) {
            const double2 ea = load_eta(stage, i0);
            if (nc > 0) {
                const uint32_t sa = cs.slot0 + stage * (RING_OPS * 512u);
                const RowPair<FAMILY> rp[1] = {RowPair<FAMILY>(lds2(sa + 512u), ea, lds2(sa + 1024u))};
                score_pairs<FAMILY, 1>(rp, nc, cmask, s_dl, d.inv_sd, tab, sacc, lane, bE, bX, nearmask, smax);
                nrows += 2;
            }
        }
        cs.issue(T + RING_D * cs.W, (int)stage, lane);
        stage = (stage + 2) & (RING_D - 1);
    }
    cp_async_wait<0>();
    if (n & 1) {  // odd last row of the matrix: one lane of one worker, scalar
        const int64_t t = n - 1;
        const long long Tl = t / TILE_ROWS;
        if (Tl % cs.W == cs.vw && lane == (int)((t % TILE_ROWS) >> 1)) {
            double e = __ldcg(eta + t);
            if (cj >= 0) { e = eta_shift(e, __ldg(cs.xc + t), cdelta); eta[t] = e; }
            const double yy = __ldg(cs.y + t), xx = __ldg(cs.xj + t);
            for (int k = 0; k < nc; ++k) sacc[k * 32 + lane] += row_term<FAMILY>(yy, eta_shift(e, xx, s_dl[k]), d.inv_sd, tab);
        }
    }
    if (FAMILY == CGG_BINOMIAL && cmask && nrows) {
        // The pre-filter's bound is against the smooth -softplus; an exact evaluation follows R's log(1 - p) form, which
        // deviates by at most 2^-54 (1 + e^s) per row for s <= 30 (rform_log1p_rho; above the clamp it is a constant).
        // Folded into the sum the decider multiplies by (2^-22 + kappa) = 7.15e-7: 5.6e-17 / 7.15e-7, rounded up.
        bE += ((float)nrows + smax) * 8.2e-11f;
    }
}

// One warp, one chain, one JET pass (cgg_jet.cuh): applies the pending eta update like any pass and accumulates
// the exact log-likelihood at the committed eta plus the derivative moments along column j in registers.
template <int FAMILY, bool FULL>
__device__ __forceinline__ void jet_tile(const ChainStream &cs, double cdelta, double cscale, double inv_sd, int stage, double *eta_i,
                                         const double2 *tab, double (&m)[NV], unsigned &risk, unsigned &rows) {
    rows += 2;
    const uint32_t s = cs.slot0 + (uint32_t)stage * (RING_OPS * 512u);
    double2 e = lds2(s);
    if (cs.cj >= 0) {
        const double2 cv = lds2(s + 1536u);
        e.x = eta_shift(e.x, cv.x, cdelta);
        e.y = eta_shift(e.y, cv.y, cdelta);
        *reinterpret_cast<double2 *>(eta_i) = e;
    }
    double2 xs = lds2(s + 1024u);
    xs.x *= cscale; xs.y *= cscale;           // power of two: exact
    JetRow<FAMILY>::template add2<FULL>(lds2(s + 512u), e, xs, inv_sd, tab, m, risk);
}

template <int FAMILY, bool FULL>
__device__ __forceinline__ void warp_pass_jet(const Dev &d, const ChainStream &cs, double cdelta, double cscale, int lane,
                                              const double2 *tab, bool prefetched, double (&m)[NV]) {
    const int cj = cs.cj;
    const int64_t n = cs.n;
    double *eta = cs.eta;
    if (!prefetched) cs.prologue(lane);
#pragma unroll
    for (int k = 0; k < NV; ++k) m[k] = 0.0;
    unsigned risk = 0;     // running max of a per-row integer key (JetRow): compared with the family's threshold at the end
    unsigned rows = 0;     // rows this lane scored
    // Running pointers of the tile being ISSUED (RING_D - 1 tiles ahead of the one being scored), advanced by one
    // stride per iteration: four 64-bit adds instead of re-deriving four addresses from the chain/column indices.
    {
        static_assert((RING_D & (RING_D - 1)) == 0, "ring depth must be a power of two");
        unsigned stage = 0;
        const int64_t step = cs.W * TILE_ROWS;
        const int64_t i0 = cs.vw * TILE_ROWS + 2 * lane;
        const double *pe = eta + i0 + (RING_D - 1) * step, *py = cs.y + i0 + (RING_D - 1) * step;     // not const: issue_next advances them
        const double *px = cs.xj + i0 + (RING_D - 1) * step, *pc = cs.xc + i0 + (RING_D - 1) * step;
        const double *const pe_last = eta + (n - 1);                   // a pair at p is inside the matrix iff p < pe_last
        const double *const pe_end = eta + cs.n_tiles * TILE_ROWS + 2 * lane + (RING_D - 1) * step;
        const uint32_t sbase = cs.slot0;
        auto issue_next = [&](unsigned st) {       // the tile at the running pointers goes to ring stage st
            const bool ok = pe < pe_last;          // (predicated copies, no branch: see cp_async16_if)
            const uint32_t sa = sbase + st * (RING_OPS * 512u);
            cp_async16_if(ok, sa, pe);
            cp_async16_if(ok, sa + 512u, py); cp_async16_if(ok, sa + 1024u, px);
            cp_async16_if(ok && cj >= 0, sa + 1536u, pc);
            cp_async_commit();
            pe += step; py += step; px += step; pc += step;
        };
#if CGG_JET_TPI == 2
        // Two tiles (four rows per lane) per iteration: with 2 warps per scheduler the loop is bound by the dependency
        // latency of a row's exp -> reciprocal -> moments chain, and the second tile's chain fills the gaps.
        static_assert(RING_D >= 4, "two tiles per iteration need a ring of at least 4 stages");
        for (; pe < pe_end; ) {
            issue_next((stage + RING_D - 1) & (RING_D - 1));
            cp_async_wait<RING_D - 2>();
            double *e0 = const_cast<double *>(pe) - RING_D * step, *e1 = e0 + step;
            if (e1 < pe_last) {
                jet_tile<FAMILY, FULL>(cs, cdelta, cscale, d.inv_sd, (int)stage, e0, tab, m, risk, rows);
                jet_tile<FAMILY, FULL>(cs, cdelta, cscale, d.inv_sd, (int)((stage + 1) & (RING_D - 1)), e1, tab, m, risk, rows);
            } else if (e0 < pe_last) {
                jet_tile<FAMILY, FULL>(cs, cdelta, cscale, d.inv_sd, (int)stage, e0, tab, m, risk, rows);
            }
            issue_next(stage);
            stage = (stage + 2) & (RING_D - 1);
        }
#else
        for (; pe < pe_end; ) {
            issue_next((stage + RING_D - 1) & (RING_D - 1));
            cp_async_wait<RING_D - 1>();
            double *ecur = const_cast<double *>(pe) - RING_D * step;
            if (ecur < pe_last) jet_tile<FAMILY, FULL>(cs, cdelta, cscale, d.inv_sd, (int)stage, ecur, tab, m, risk, rows);
            stage = (stage + 1) & (RING_D - 1);
        }
#endif
    }
    cp_async_wait<0>();
    if (n & 1) {  // odd last row of the matrix: one lane of one worker, scalar
        const int64_t t = n - 1;
        const long long Tl = t / TILE_ROWS;
        if (Tl % cs.W == cs.vw && lane == (int)((t % TILE_ROWS) >> 1)) {
            double e = __ldcg(eta + t);
            if (cj >= 0) { e = eta_shift(e, __ldg(cs.xc + t), cdelta); eta[t] = e; }
            JetRow<FAMILY>::template add1<FULL>(__ldg(cs.y + t), e, __ldg(cs.xj + t) * cscale, d.inv_sd, tab, m, risk);
            rows += 1;
        }
    }
    if (FAMILY != CGG_GAUSSIAN) m[9] = (risk >= JetRow<FAMILY>::RISK_KEY) ? 1.0 : 0.0;
    if (FAMILY == CGG_BINOMIAL) m[8] = rform_noise_sum(risk, rows);
    if (FAMILY == CGG_BINOMIAL && !FULL) jet_light_pack(m);
}

// One warp, TWO chains that are at the same coordinate (same column j, same pending column), one jet pass: y, X_j and
// X_commit tiles are staged once and serve both chains (the chains of a GPU run in lock-step in the stationary regime),
// the per-tile loop overhead and the pipeline fill are paid once, and the two chains' rows are independent work for the
// scheduler.  Ring stage layout: [eta_A, eta_B, y, X_j, X_commit] x 512 B.
// The X-column cache.  The tiles a warp owns are the same rows for every chain and every column (when the cache is on the
// tile -> warp map has no per-chain rotation), and its share of one column is small: n = 1e6 rows over 1176 warps is 14
// tiles = 7 KB.  Every warp therefore keeps TWO columns of its rows in shared memory, tagged with their column index:
// the column being sampled (filled by the first pair pass that walks it, served to the other pairs of the GPU) and the
// previous one, which is the next passes' X_commit.  In the lock-step stationary regime a column is then read from L2
// ONCE per warp instead of 8 times (4 pair passes as X_j, 4 as X_commit), which matters because with eta pinned in L2 the
// kernel is bound by L2 bandwidth (~8 TB/s measured): per pair pass 2 eta reads + 2 eta writes remain, instead of 5 reads
// + 2 writes.  Private to the warp: a lane only ever reads back the 16-byte slots it filled itself, so -- like the
// staging ring -- it needs no barrier.  A miss (chains out of step, an exact pass in between) just streams as before.
struct ColCache {
    uint32_t base;             // shared-space address of this warp's cache + this lane's 16 bytes; slot s, tile t at base + (s * cap + t) * 512
    int cap;                   // tiles per slot; 0: the cache is off
    int tag0, tag1;            // column held (completely) by slot 0 / slot 1; -1: none
    int fill_col, fill_slot;   // a fill in progress (begun by a prologue that was issued early); -1: none
};
struct PairStream {      // running pointers of a pair pass: the tile to be issued next
    const double *pa, *pb, *py, *px, *pc;
    const double *pa_last, *pa_end;
    double *etaA, *etaB;
    const double *xj, *xc;
    long long vw;
    int64_t step;
    int cj, j;
    uint32_t sbase;
    uint32_t xj_slot, xc_slot; // cache slots serving X_j / X_commit (this lane's column)
    int sj;                    // cache slot index of X_j (-1: staged through the ring)
    int t_issue;               // index (within this warp's tiles) of the tile to be issued next
    bool need_y, xj_ring, xc_ring, fillJ;
    __device__ __forceinline__ PairStream(const Dev &d, int c0, const double *cwA, long long wid, long long W, int lane, uint32_t ring, ColCache &cc, bool with_y) {
        const long long w0 = __double_as_longlong(cwA[0]), w1 = __double_as_longlong(cwA[1]);
        j = (int)(w0 & 0xffffffffLL);
        cj = (int)(w1 & 0xffffffffLL);
        etaA = d.eta + (int64_t)c0 * d.lde; etaB = etaA + d.lde;
        xj = d.X + (int64_t)j * d.ldx; xc = d.X + (int64_t)(cj < 0 ? 0 : cj) * d.ldx;
        vw = d.colcache ? wid : (wid + (long long)(c0 & ~1) * (W / d.C)) % W;      // == ChainStream's map of c0 and of c0 + 1 (d.pair is on)
        sbase = ring + (uint32_t)lane * 16u;
        step = W * TILE_ROWS;
        const int64_t i0 = vw * TILE_ROWS + 2 * lane;
        pa = etaA + i0; pb = etaB + i0; py = d.y + i0; px = xj + i0; pc = xc + i0;
        pa_last = etaA + (d.n - 1);
        pa_end = etaA + d.n_tiles * TILE_ROWS + 2 * lane + (RING_D - 1) * step;
        need_y = with_y; t_issue = 0;
        // ---- where X_j and X_commit come from
        sj = -1; fillJ = false; xj_ring = true; xc_ring = cj >= 0; xj_slot = xc_slot = 0;
        if (cc.cap > 0) {
            xj_ring = false;
            if (cc.tag0 == j) sj = 0;
            else if (cc.tag1 == j) sj = 1;
            else if (cc.fill_col == j) { sj = cc.fill_slot; fillJ = true; }
            else {                                     // miss: fill the slot that does not hold the commit column
                sj = (cj >= 0 && cc.tag0 == cj) ? 1 : 0;
                if (sj == 0) cc.tag0 = -1; else cc.tag1 = -1;
                cc.fill_col = j; cc.fill_slot = sj; fillJ = true;
            }
            xj_slot = cc.base + (uint32_t)(sj * cc.cap) * 512u;
            if (cj >= 0) {
                const int sc = (cc.tag0 == cj) ? 0 : ((cc.tag1 == cj) ? 1 : -1);
                if (sc >= 0) { xc_ring = false; xc_slot = cc.base + (uint32_t)(sc * cc.cap) * 512u; }
            }
        }
    }
    __device__ __forceinline__ void issue_next(unsigned st) {
        if (pa < pa_last) {
            const uint32_t sa = sbase + st * (RING_OPS * 512u);
            cp_async16(sa, pa); cp_async16(sa + 512u, pb);
            if (need_y) cp_async16(sa + 1024u, py);
            if (xj_ring) cp_async16(sa + 1536u, px);
            else if (fillJ) cp_async16(xj_slot + (uint32_t)t_issue * 512u, px);
            if (xc_ring) cp_async16(sa + 2048u, pc);
        }
        cp_async_commit();
        pa += step; pb += step; py += step; px += step; pc += step;
        ++t_issue;
    }
    // first RING_D - 1 tiles: may be issued early (cross-pair prefetch); `skip`: they already were
    __device__ __forceinline__ void prologue(bool skip) {
        if (skip) { const int64_t adv = (RING_D - 1) * step; pa += adv; pb += adv; py += adv; px += adv; pc += adv; t_issue += RING_D - 1; return; }
#pragma unroll
        for (int s = 0; s < RING_D - 1; ++s) issue_next((unsigned)s);
    }
    // the pass is over (all copies have landed): a filled column becomes a cached one
    __device__ __forceinline__ void finish(ColCache &cc) const {
        if (fillJ) { if (sj == 0) cc.tag0 = j; else cc.tag1 = j; cc.fill_col = -1; }
    }
};

template <int FAMILY, bool FULL, class EARLY, class AFTER>
__device__ __forceinline__ void warp_pass_jet2(const Dev &d, int c0, const double *cwA, const double *cwB, long long wid, long long W,
                                               int lane, uint32_t ring, const double2 *tab, bool prefetched, ColCache &cc, EARLY &&after_prologue,
                                               AFTER &&after_tiles, double (&mA)[NV], double (&mB)[NV]) {
    const double cdA = cwA[2], cdB = cwB[2], cscale = cwA[CTL_WORDS - 1];
    const int64_t n = d.n;
    constexpr bool WITH_Y = FAMILY != CGG_BINOMIAL || FULL;        // a binomial light pass never looks at y (C1 carries it)
    PairStream ps(d, c0, cwA, wid, W, lane, ring, cc, WITH_Y);
    const int cj = ps.cj;
    double *etaA = ps.etaA, *etaB = ps.etaB;
    constexpr uint32_t STAGE = RING_OPS * 512u;
    static_assert(RING_OPS >= 5, "a pair pass stages up to five operands");
#pragma unroll
    for (int k = 0; k < NV; ++k) { mA[k] = 0.0; mB[k] = 0.0; }
    unsigned riskA = 0, riskB = 0, rows = 0;
    ps.prologue(prefetched);
    after_prologue();     // the first tiles are on their way: a good moment for a (blocking) look at the next pair's decision
    unsigned stage = 0;
    uint32_t t_off = 0;   // byte offset of the tile being scored inside a cache slot
    auto score_tile = [&](unsigned st, int64_t off) {
        const uint32_t s = ps.sbase + st * STAGE;
        double2 ea = lds2(s), eb = lds2(s + 512u);
        if (cj >= 0) {
            const double2 cv = ps.xc_ring ? lds2(s + 2048u) : lds2(ps.xc_slot + t_off);
            ea.x = eta_shift(ea.x, cv.x, cdA); ea.y = eta_shift(ea.y, cv.y, cdA);
            eb.x = eta_shift(eb.x, cv.x, cdB); eb.y = eta_shift(eb.y, cv.y, cdB);
#ifndef CGG_DIAG_NOSTORE      // (diagnostic builds only: what does the write-back cost?  results are wrong without it)
            *reinterpret_cast<double2 *>(etaA + off) = ea;
            *reinterpret_cast<double2 *>(etaB + off) = eb;
#endif
        }
        double2 yy = make_double2(0.0, 0.0);
        if (WITH_Y) yy = lds2(s + 1024u);
        double2 xs = ps.xj_ring ? lds2(s + 1536u) : lds2(ps.xj_slot + t_off);
        xs.x *= cscale; xs.y *= cscale;
        rows += 2;
#ifdef CGG_DIAG_NOMATH        // (diagnostic builds only: what does the row math cost?)
        mA[1] += ea.x + ea.y + xs.x + yy.x; mB[1] += eb.x + eb.y + xs.y + yy.y;
#else
        JetRow<FAMILY>::template add2<FULL>(yy, ea, xs, d.inv_sd, tab, mA, riskA);
        JetRow<FAMILY>::template add2<FULL>(yy, eb, xs, d.inv_sd, tab, mB, riskB);
#endif
        t_off += 512u;
    };
#if CGG_PAIR_TPI == 2
    for (; ps.pa < ps.pa_end; ) {
        ps.issue_next((stage + RING_D - 1) & (RING_D - 1));
        cp_async_wait<RING_D - 2>();
        const int64_t off0 = (ps.pa - etaA) - RING_D * ps.step, off1 = off0 + ps.step;
        if (off1 + 1 < n) {
            score_tile(stage, off0);
            score_tile((stage + 1) & (RING_D - 1), off1);
        } else if (off0 + 1 < n) {
            score_tile(stage, off0);
        }
        ps.issue_next(stage);
        stage = (stage + 2) & (RING_D - 1);
    }
#else
    for (; ps.pa < ps.pa_end; ) {
        ps.issue_next((stage + RING_D - 1) & (RING_D - 1));
        cp_async_wait<RING_D - 1>();
        const int64_t off = (ps.pa - etaA) - RING_D * ps.step;        // row index of this lane's pair in the tile being scored
        if (off + 1 < n) score_tile(stage, off);
        stage = (stage + 1) & (RING_D - 1);
    }
#endif
    cp_async_wait<0>();
    ps.finish(cc);
    if (n & 1) {  // odd last row of the matrix: one lane of one worker, scalar
        const int64_t t = n - 1;
        const long long Tl = t / TILE_ROWS;
        if (Tl % W == ps.vw && lane == (int)((t % TILE_ROWS) >> 1)) {
            double ea = __ldcg(etaA + t), eb = __ldcg(etaB + t);
            if (cj >= 0) {
                const double cv = __ldg(ps.xc + t);
                ea = eta_shift(ea, cv, cdA); eb = eta_shift(eb, cv, cdB);
                etaA[t] = ea; etaB[t] = eb;
            }
            const double yy = __ldg(d.y + t), xx = __ldg(ps.xj + t) * cscale;
            JetRow<FAMILY>::template add1<FULL>(yy, ea, xx, d.inv_sd, tab, mA, riskA);
            JetRow<FAMILY>::template add1<FULL>(yy, eb, xx, d.inv_sd, tab, mB, riskB);
            rows += 1;
        }
    }
    after_tiles();     // the ring is free: the caller may already request the next pair's first tiles
    if (FAMILY != CGG_GAUSSIAN) {
        mA[9] = (riskA >= JetRow<FAMILY>::RISK_KEY) ? 1.0 : 0.0;
        mB[9] = (riskB >= JetRow<FAMILY>::RISK_KEY) ? 1.0 : 0.0;
    }
    if (FAMILY == CGG_BINOMIAL) { mA[8] = rform_noise_sum(riskA, rows); mB[8] = rform_noise_sum(riskB, rows); }
    if (FAMILY == CGG_BINOMIAL && !FULL) { jet_light_pack(mA); jet_light_pack(mB); }
    constexpr int NVD = (FAMILY == CGG_BINOMIAL && !FULL) ? JET_NVL : NV;      // values this pass delivers
#pragma unroll
    for (int k = 0; k < NVD; ++k) { mA[k] = warp_sum(mA[k]); mB[k] = warp_sum(mB[k]); }
}

// ---------------------------------------------------------------------------------------------
// GROUP passes: NCH chains (c0 .. c0 + NCH - 1) at the same coordinate share one walk over the rows -- the pair pass
// carried further.  What a walk costs per TILE (ring bookkeeping, branches, the reads of X_j and X_commit, the
// look-ahead, the reductions' and the delivery's fixed part, the wait for the slowest CTA) is paid once for NCH chains
// instead of once per pair: of the pair loop's 242 instructions per tile about 100 are such per-tile work.
// A group pass is the specialisation for the steady state and ONLY for it: the X-column cache is on, X_commit is held
// by the cache (or there is no pending update) and X_j is held or being filled by this very pass -- so the loop has
// no operand-source branches, and a ring stage is [eta_0 .. eta_{NCH-1}, y] x 512 B.  Anything else runs as pair or
// single passes.  Which kind of walk a WARP uses for a pass is its own business: every warp owns the same rows in
// every kind of pass (with the cache on the tile -> warp map has no rotation) and delivers once per chain and pass.
__device__ __forceinline__ bool group_cache_ok(const ColCache &cc, int cj) {
    return cc.cap > 0 && (cj < 0 || cc.tag0 == cj || cc.tag1 == cj);
}
template <int NCH>
struct GroupStream {
    const double *pe;          // chain c0's eta at this lane's pair of the tile to be issued next
    const double *pe_last, *pe_end;
    double *eta0;              // chain c0's eta; chain k's is eta0 + k * lde
    const double *y0, *xj0, *xc0;
    int64_t lde, step;
    long long vw;
    int cj, j, sj, t_issue;
    uint32_t sbase, xj_slot, xc_slot;
    bool need_y, fillJ;
    __device__ __forceinline__ GroupStream(const Dev &d, int c0, const double *cw, long long wid, long long W, int lane, uint32_t ring,
                                           ColCache &cc, bool with_y) {
        const long long w0 = __double_as_longlong(cw[0]), w1 = __double_as_longlong(cw[1]);
        j = (int)(w0 & 0xffffffffLL);
        cj = (int)(w1 & 0xffffffffLL);
        lde = d.lde;
        eta0 = d.eta + (int64_t)c0 * d.lde;
        y0 = d.y; xj0 = d.X + (int64_t)j * d.ldx; xc0 = d.X + (int64_t)(cj < 0 ? 0 : cj) * d.ldx;
        vw = wid;                                   // the cache is on: no rotation (== ChainStream's and PairStream's map)
        sbase = ring + (uint32_t)lane * 16u;
        step = W * TILE_ROWS;
        const int64_t i0 = vw * TILE_ROWS + 2 * lane;
        pe = eta0 + i0;
        pe_last = eta0 + (d.n - 1);
        pe_end = eta0 + d.n_tiles * TILE_ROWS + 2 * lane + (RING_D - 1) * step;
        need_y = with_y; t_issue = 0;
        // ---- X_j: held, being filled (a prologue issued early began it), or a miss that this pass fills
        fillJ = false;
        if (cc.tag0 == j) sj = 0;
        else if (cc.tag1 == j) sj = 1;
        else if (cc.fill_col == j) { sj = cc.fill_slot; fillJ = true; }
        else {
            sj = (cj >= 0 && cc.tag0 == cj) ? 1 : 0;
            if (sj == 0) cc.tag0 = -1; else cc.tag1 = -1;
            cc.fill_col = j; cc.fill_slot = sj; fillJ = true;
        }
        xj_slot = cc.base + (uint32_t)(sj * cc.cap) * 512u;
        xc_slot = cc.base + (uint32_t)(((cj >= 0 && cc.tag0 == cj) ? 0 : 1) * cc.cap) * 512u;    // (the caller checked group_cache_ok)
    }
    __device__ __forceinline__ void issue_next(unsigned st) {
        const bool ok = pe < pe_last;
        const uint32_t sa = sbase + st * (RING_OPS * 512u);
#pragma unroll
        for (int k = 0; k < NCH; ++k) cp_async16_if(ok, sa + (uint32_t)k * 512u, pe + k * lde);
        if (need_y) cp_async16_if(ok, sa + (uint32_t)NCH * 512u, y0 + (pe - eta0));           // (need_y is a compile-time constant of the pass)
        cp_async16_if(ok && fillJ, xj_slot + (uint32_t)t_issue * 512u, xj0 + (pe - eta0));
        cp_async_commit();
        pe += step;
        ++t_issue;
    }
    __device__ __forceinline__ void prologue(bool skip) {
        if (skip) { pe += (RING_D - 1) * step; t_issue += RING_D - 1; return; }
#pragma unroll
        for (int s = 0; s < RING_D - 1; ++s) issue_next((unsigned)s);
    }
    __device__ __forceinline__ void finish(ColCache &cc) const {
        if (fillJ) { if (sj == 0) cc.tag0 = j; else cc.tag1 = j; cc.fill_col = -1; }
    }
};

// cw0: the CTA's shared copies of the control blocks of chains c0 .. c0 + NCH - 1 (consecutive, CTL_WORDS apart), all
// of them saying the same thing (pair_batchable).  m[k]: the sums of chain c0 + k, reduced over the warp.
template <int FAMILY, bool FULL, int NCH, class EARLY, class AFTER>
__device__ __forceinline__ void warp_pass_group(const Dev &d, int c0, const double *cw0, long long wid, long long W, int lane, uint32_t ring,
                                                const double2 *tab, bool prefetched, ColCache &cc, EARLY &&after_prologue, AFTER &&after_tiles,
                                                double (&m)[NCH][NV]) {
    static_assert(NCH + 1 <= RING_OPS, "a group pass stages NCH eta tiles and y");
    constexpr bool WITH_Y = FAMILY != CGG_BINOMIAL || FULL;        // a binomial light pass never looks at y (C1 carries it)
    constexpr uint32_t STAGE = RING_OPS * 512u;
    double cd[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) cd[k] = cw0[k * CTL_WORDS + 2];
    const double cscale = cw0[CTL_WORDS - 1];
    const int64_t n = d.n;
    GroupStream<NCH> gs(d, c0, cw0, wid, W, lane, ring, cc, WITH_Y);
    const int cj = gs.cj;
    double *eta0 = gs.eta0;
    const int64_t lde = gs.lde;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
        for (int q = 0; q < NV; ++q) m[k][q] = 0.0;
    }
    unsigned risk[NCH], rows = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) risk[k] = 0;
    gs.prologue(prefetched);
    after_prologue();
    unsigned stage = 0;
    uint32_t t_off = 0;   // byte offset of the tile being scored inside a cache slot
    // chain c0's eta at this lane's pair of the tile being scored (RING_D - 1 tiles behind the one being issued).  The pending
    // update is applied unconditionally: without one, commit_delta is 0 and X_commit reads as 0, so eta + 0 * 0 is written back
    // unchanged (the first pass of a launch only) -- no branch in the loop.
    double *ps = const_cast<double *>(gs.pe) - (RING_D - 1) * gs.step;      // (the prologue, issued now or earlier, is RING_D - 1 tiles ahead)
    const int64_t lde_b = lde;
    for (; gs.pe < gs.pe_end; ) {
        gs.issue_next((stage + RING_D - 1) & (RING_D - 1));
        cp_async_wait<RING_D - 1>();
        if (ps < gs.pe_last) {
            const uint32_t s = gs.sbase + stage * STAGE;
            double2 xs = lds2(gs.xj_slot + t_off);
            xs.x *= cscale; xs.y *= cscale;
            double2 cv = make_double2(0.0, 0.0), yy = make_double2(0.0, 0.0);
            if (cj >= 0) cv = lds2(gs.xc_slot + t_off);
            if (WITH_Y) yy = lds2(s + (uint32_t)NCH * 512u);
            rows += 2;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                double2 e = lds2(s + (uint32_t)k * 512u);
                e.x = eta_shift(e.x, cv.x, cd[k]); e.y = eta_shift(e.y, cv.y, cd[k]);
                *reinterpret_cast<double2 *>(ps + k * lde_b) = e;
                JetRow<FAMILY>::template add2<FULL>(yy, e, xs, d.inv_sd, tab, m[k], risk[k]);
            }
        }
        ps += gs.step;
        t_off += 512u;
        stage = (stage + 1) & (RING_D - 1);
    }
    cp_async_wait<0>();
    gs.finish(cc);
    if (n & 1) {  // odd last row of the matrix: one lane of one worker, scalar
        const int64_t t = n - 1;
        const long long Tl = t / TILE_ROWS;
        if (Tl % W == gs.vw && lane == (int)((t % TILE_ROWS) >> 1)) {
            const double cv = (cj >= 0) ? __ldg(gs.xc0 + t) : 0.0;
            const double yy = __ldg(d.y + t), xx = __ldg(gs.xj0 + t) * cscale;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                double e = __ldcg(eta0 + k * lde + t);
                if (cj >= 0) { e = eta_shift(e, cv, cd[k]); eta0[k * lde + t] = e; }
                JetRow<FAMILY>::template add1<FULL>(yy, e, xx, d.inv_sd, tab, m[k], risk[k]);
            }
            rows += 1;
        }
    }
    after_tiles();     // the ring is free: the caller may already request the next group's first tiles
    constexpr int NVD = (FAMILY == CGG_BINOMIAL && !FULL) ? JET_NVL : NV;      // values this pass delivers
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        if (FAMILY != CGG_GAUSSIAN) m[k][9] = (risk[k] >= JetRow<FAMILY>::RISK_KEY) ? 1.0 : 0.0;
        if (FAMILY == CGG_BINOMIAL) m[k][8] = rform_noise_sum(risk[k], rows);
        if (FAMILY == CGG_BINOMIAL && !FULL) jet_light_pack(m[k]);
#pragma unroll
        for (int q = 0; q < NVD; ++q) m[k][q] = warp_sum(m[k][q]);
    }
}

// Two control blocks describe passes that can share one walk over the rows: both jet passes of the same kind on the
// same column with the same pending column.
__device__ __forceinline__ bool pair_batchable(const double *cwA, const double *cwB) {
    const long long a0 = __double_as_longlong(cwA[0]), a1 = __double_as_longlong(cwA[1]);
    const long long b0 = __double_as_longlong(cwB[0]), b1 = __double_as_longlong(cwB[1]);
    const unsigned ma = (unsigned)(a1 >> 32);
    return a0 == b0 && a1 == b1 && (ma & JET_BIT) && (int)(a0 & 0xffffffffLL) >= 0;
}

// A worker's whole contribution to one pass of chain c: read the control block, stream the rows and
// return the warp's partial sums (identical in every lane): acc[0..nc) candidate sums and, when the
// pre-filter is active, acc[nc], acc[nc+1] = sum |eta|, sum |x| over the warp's rows.  Return value: -1
// chain finished, otherwise the number of values delivered (0 when the pass was idle or commit-only).
// next_cw: the control block of the chain this warp will stream next if its decision is already published
// (else nullptr); its first ring stages are then issued as soon as this chain's tiles are consumed, so the
// loads are in flight during the reduction and delivery below.  `prefetched` is updated accordingly.
// No look-ahead (stepwise driver: every launch starts from scratch).
struct NoLookAhead { __device__ __forceinline__ const double *poll(const Dev &, int) { return nullptr; } };

template <int FAMILY, class LA>
__device__ __forceinline__ int worker_pass(const Dev &d, int c, const double *cw /* the CTA's shared copy of ctl[c] */,
                                           long long wid, long long W, int lane, uint32_t ring, const double2 *tab,
                                           double *sacc, double (&acc)[NV], int &j_out, bool &prefetched,
                                           int next_c, LA *la, long long *t_tiles = nullptr) {
    const long long w0 = __double_as_longlong(cw[0]), w1 = __double_as_longlong(cw[1]);
    const int j = (int)(w0 & 0xffffffffLL), nc = (int)(w0 >> 32), cj = (int)(w1 & 0xffffffffLL);
    const unsigned cmask = (unsigned)(w1 >> 32);
    j_out = j;
    const bool was_prefetched = prefetched;
    prefetched = false;
    if (j < 0) return -1;
    if (cmask & JET_BIT) {
        {
            const long long t0 = t_tiles ? clock64() : 0;
            const ChainStream cs(d, c, cw, wid, W, lane, ring);
            if (FAMILY != CGG_BINOMIAL || (cmask & JET_FULL)) warp_pass_jet<FAMILY, true>(d, cs, cw[2], cw[CTL_WORDS - 1], lane, tab, was_prefetched, acc);
            else warp_pass_jet<FAMILY, false>(d, cs, cw[2], cw[CTL_WORDS - 1], lane, tab, was_prefetched, acc);
            if (t_tiles) *t_tiles += clock64() - t0;
        }
        const int nvd = jet_nvals(FAMILY, !(cmask & JET_FULL));
        if (const double *next_cw = la->poll(d, lane)) {
            const ChainStream ns(d, next_c, next_cw, wid, W, lane, ring);
            const long long nw0 = __double_as_longlong(next_cw[0]);
            const int nj = (int)(nw0 & 0xffffffffLL);
            if (nj >= 0 && (ns.need_yx || ns.cj >= 0)) { ns.prologue(lane); prefetched = true; }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) if (k < nvd) acc[k] = warp_sum(acc[k]);
        return nvd;
    }
    if (nc == 0 && cj < 0) return 0;
    float bE = 0.0f, bX = 0.0f;
    unsigned nearmask = 0;
    {
        const long long t0 = t_tiles ? clock64() : 0;
        const ChainStream cs(d, c, cw, wid, W, lane, ring);
        warp_pass_chain<FAMILY>(d, cs, cw[2], cw + 3, lane, tab, sacc, was_prefetched, cmask, bE, bX, nearmask);
        if (t_tiles) *t_tiles += clock64() - t0;
    }
    if (const double *next_cw = la->poll(d, lane)) {
        const ChainStream ns(d, next_c, next_cw, wid, W, lane, ring);
        const long long nw0 = __double_as_longlong(next_cw[0]);
        const int nj = (int)(nw0 & 0xffffffffLL);
        if (nj >= 0 && (ns.need_yx || ns.cj >= 0)) { ns.prologue(lane); prefetched = true; }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < nc) acc[k] = warp_sum(sacc[k * 32 + lane]);
    if (!cmask) return nc;
    const double sE = warp_sum((double)bE), sX = warp_sum((double)bX);
#pragma unroll
    for (int k = 0; k < NV; ++k) { if (k == nc) acc[k] = sE; if (k == nc + 1) acc[k] = sX; }
    nearmask = __reduce_or_sync(0xffffffffu, nearmask);
    if (nearmask) {                // ... and as a NaN bound sum, which travels with the values themselves (slot delivery)
#pragma unroll
        for (int k = 0; k < NV; ++k) if (k == nc) acc[k] = NAN;
    }
    if (nearmask && lane == 0) {   // a row sat on the |eta| = 30 clamp discontinuity: the bound does not hold
        for (int k = 0; k < nc; ++k)
            if ((nearmask >> k) & 1u) atomicOr(&d.acc[c * NV + k].flags, 8u);
        fence_gpu();
    }
    return nc + 2;
}

// Shared memory of a sweep CTA: the warps' staging rings, per-chain slots of the non-blocking CTA-level
// reduction and the CTA's cached view of the chains' version flags.
struct CtaShared {                       // views into dynamic shared memory, sized by the chain count
    double *part;                        // [C][NWARPS][NV] warp partial sums of the pass in flight
    double *ctl;                         // [C][CTL_WORDS] the control block that goes with ver[c] (one L2 fetch per CTA)
    unsigned long long *ver;             // [C] last version of chain c seen by this CTA
    int *cnt;                            // [C] warps of this CTA that delivered their partials
    int *lock;                           // [C] elected poller of the chain's version flag
    double *sacc;                        // [NWARPS][KMAX][32] per-lane running sums of the chain pass in flight
    uint32_t ring0;                      // shared-space address of warp 0's ring
    uint32_t cache0;                     // shared-space address of warp 0's X-column cache (behind everything else)
    __device__ __forceinline__ CtaShared(unsigned char *base, int C) {
        ring0 = (uint32_t)__cvta_generic_to_shared(base);
        cache0 = ring0 + (uint32_t)((bytes(C) + 15) / 16 * 16);
        sacc = reinterpret_cast<double *>(base + NWARPS * RING_BYTES_PER_WARP);
        part = sacc + NWARPS * KMAX * 32;
        ctl = part + (size_t)C * NWARPS * NV;
        ver = reinterpret_cast<unsigned long long *>(ctl + (size_t)C * CTL_WORDS);
        cnt = reinterpret_cast<int *>(ver + C);
        lock = cnt + C;
    }
    __host__ __device__ static size_t bytes(int C) {
        return (size_t)NWARPS * RING_BYTES_PER_WARP + sizeof(double) * NWARPS * KMAX * 32 +
               (size_t)C * (sizeof(double) * (NWARPS * NV + CTL_WORDS) + sizeof(unsigned long long) + 2 * sizeof(int));
    }
};

// Persistent driver: at the END of a chain's pass (tiles consumed, sums not yet reduced) look whether the next chain's
// decision is published; if so bring its control block into shared memory (one elected warp per CTA, everybody else
// only reads shared memory) so that its first tiles can be requested before this chain's sums are reduced and
// delivered, and nobody has to fetch the block from global memory at the top of the next pass.  Decisions are
// published roughly one pass before they are needed, so a look at the start of the pass would be too early.
// The CTA's shared view of a chain's version is read by ONE lane and the verdict shared with a warp vote: every branch
// of the hand-shake below is then taken by all 32 lanes or by none.  A per-lane `volatile` read is not good enough: the
// lanes of a warp are only guaranteed to be converged at *_sync primitives, another warp updates the word at any time,
// and lanes that read it a few cycles apart took different branches -- one lane then met its warp's next __shfl_sync at
// a different call site (seen on the GPU: a version "read" as the halves of two partial sums).
__device__ __forceinline__ unsigned long long ver_read(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ bool ver_below(const unsigned long long *p, unsigned long long round, int lane) {     // shared version < round ?
    return __any_sync(0xffffffffu, lane == 0 && ver_read(p) < round);
}
__device__ __forceinline__ bool ver_is(const unsigned long long *p, unsigned long long round, int lane) {        // shared version == round ?
    return __any_sync(0xffffffffu, lane == 0 && ver_read(p) == round);
}

struct LookAhead {
    CtaShared *sh; const ChainSync *sync; const Ctl *ctl;
    int nxt; unsigned long long nround;
    long long n_look, n_ok;
    __device__ __forceinline__ const double *poll(const Dev &d, int lane) {
        if (nxt < 0) return nullptr;
        volatile unsigned long long *sv = &sh->ver[nxt];
        if (ver_below(&sh->ver[nxt], nround, lane)) {
            int got = 0;
            if (lane == 0) got = (atomicCAS_block(&sh->lock[nxt], 0, 1) == 0);
            got = __shfl_sync(0xffffffffu, got, 0);
            if (got) {
                unsigned long long v = 0;
                if (lane == 0) v = ld_acquire_u64(&sync[nxt].version);
                v = __shfl_sync(0xffffffffu, v, 0);
                ++n_look; if (v >= nround) ++n_ok;
                if (__any_sync(0xffffffffu, lane == 0 && v >= nround && v > ver_read(&sh->ver[nxt]))) {
                    if (lane < CTL_WORDS) sh->ctl[nxt * CTL_WORDS + lane] = __ldcg(reinterpret_cast<const double *>(ctl + nxt) + lane);
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); *sv = v; }
                }
                if (lane == 0) { __threadfence_block(); atomicExch_block(&sh->lock[nxt], 0); }
                __syncwarp();
            }
        }
        // the shared control block of `nxt` belongs to version ver[nxt]: usable only if that is exactly the pass we will run
        return ver_is(&sh->ver[nxt], nround, lane) ? sh->ctl + nxt * CTL_WORDS : nullptr;
    }
};

// The same for a PAIR of chains (c2, c2 + 1) with one election and two concurrent round trips (both version flags, then
// both control blocks).  True when both shared control blocks are exactly those of pass `nround`.
__device__ __forceinline__ bool pair_lookahead(const Dev &d, CtaShared &sh, int c2, unsigned long long nround, int lane) {
    // lane k < 2 looks after chain c2 + k: its shared version, its global version, and -- when that has moved on -- the
    // publication of the fetched block; the other lanes learn what they have to know through votes
    unsigned long long sv = (lane < 2) ? ver_read(&sh.ver[c2 + lane]) : ~0ULL;
    if (__any_sync(0xffffffffu, sv < nround)) {
        int got = 0;
        if (lane == 0) got = (atomicCAS_block(&sh.lock[c2], 0, 1) == 0);
        if (__any_sync(0xffffffffu, got != 0)) {
            unsigned long long v = 0;
            if (lane < 2) { v = ld_acquire_u64(&d.sync[c2 + lane].version); sv = ver_read(&sh.ver[c2 + lane]); }
            const bool fresh = lane < 2 && v >= nround && v > sv;         // (a stale view repeats a fetch at worst: versions only grow)
            const unsigned fm = __ballot_sync(0xffffffffu, fresh);
            if (fm) {
                const int which = lane / CTL_WORDS, word = lane % CTL_WORDS;       // lanes 0..11: chain c2, 12..23: chain c2 + 1
                if (lane < 2 * CTL_WORDS && ((fm >> which) & 1u))
                    sh.ctl[(c2 + which) * CTL_WORDS + word] = __ldcg(reinterpret_cast<const double *>(d.ctl + c2 + which) + word);
                __syncwarp();
                if (fresh) { __threadfence_block(); *reinterpret_cast<volatile unsigned long long *>(&sh.ver[c2 + lane]) = v; }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicExch_block(&sh.lock[c2], 0); }
        }
    }
    sv = (lane < 2) ? ver_read(&sh.ver[c2 + lane]) : nround;
    return __all_sync(0xffffffffu, sv == nround);
}

// Deliver a warp's partial sums.  The last warp of the CTA to deliver (returns true in all its lanes)
// has folded the CTA's NWARPS partials -- summed in warp order, so the value is reproducible -- into the
// chain's exact accumulators; the other warps return at once and move on to their next chain.
__device__ __forceinline__ bool cta_deliver(const Dev &d, CtaShared &sh, int c, int nc, int warp, int lane, int nworkers, const double (&acc)[NV]) {
    int last = 0;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k)
            if (k < nc) sh.part[((size_t)c * NWARPS + warp) * NV + k] = acc[k];
        __threadfence_block();
        last = (atomicAdd_block(&sh.cnt[c], 1) == nworkers - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return false;
    __threadfence_block();
    if (lane < nc) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w)
            if (w < nworkers) v += sh.part[((size_t)c * NWARPS + w) * NV + lane];
        acc_add(d.acc + c * NV + lane, v);
        fence_gpu();       // fences are per thread: each adding lane orders its atomics before the arrival
    }
    if (lane == 0) sh.cnt[c] = 0;
    __syncwarp();
    return true;
}

// ---------------------------------------------------------------------------------------------
// Persistent driver: how a pass's sums travel from the G worker CTAs to the chain's deciding warp.
//
// Every value of a chain has a LIMB ACCUMULATOR: four 64-bit words in global memory that only ever receive integer
// `red.add` operations (fire-and-forget atomics: nobody waits for them).  A CTA's sum v (the fold of its warps' partials)
// is split EXACTLY into four signed 32-bit-wide limbs v = p0 2^32 + p1 + p2 2^-32 + p3 2^-64 and the CTA adds, to word l,
//     2^56  +  (p_l + 2^47).
// The top 8 bits of a word therefore COUNT the CTAs that have added to it, and the low 56 bits hold the sum of their
// payloads (each payload + 2^47 is positive and G <= 255 of them stay below 2^56: no carry into the count), so each word
// validates itself: the decider reads the words of the values it expects with ONE load per lane, and when every count
// field has advanced by G since the previous pass, the differences of the low fields ARE the pass's sums.  No stamps, no
// flags, no fences, nothing is ever reset during a launch (the decider keeps the previous words; arithmetic is mod 2^64),
// and -- integer addition being associative -- the totals do not depend on arrival order: bit-reproducible.
// Resolution 2^-64 per contribution (absolute 147 x 2.7e-20: below one ulp of any sum the deciders use); magnitudes from
// 2^68 on, infinities and NaN travel as marker payloads no finite sum can reach.
// Round 1 gave every CTA its own slot per chain and value (147 x 10 stamped 16-byte entries per chain): the decider then
// needed ~50 dependent-on-nothing-but-still-slow loads per lane and spent 16 of its 27 us per decision reading them.
struct __align__(128) LimbAcc { unsigned long long w[4]; unsigned long long pad[12]; };    // one value: one L2 line
constexpr int LIMB_OFFSET_LOG2 = 47;
constexpr long long LIMB_MARK = 1LL << 45;     // marker payloads: limb 0 -MARK: a -Inf term; limb 1 +MARK: +Inf / out of range; limb 2 +MARK: NaN
__device__ __forceinline__ void limbs_split(double v, long long (&p)[4]) {
    if (!(fabs(v) < 2.9514790517935283e20 /* 2^68 */)) {
        p[0] = (v < 0.0) ? -LIMB_MARK : 0; p[1] = (v > 0.0) ? LIMB_MARK : 0; p[2] = (v != v) ? LIMB_MARK : 0; p[3] = 0;
        return;
    }
    const double h0 = trunc(v * 2.3283064365386963e-10 /* 2^-32 */);      // |h0| < 2^36
    double r = fma(-h0, 4294967296.0, v);                                   // exact: the low part of v, |r| < 2^32
    const double h1 = trunc(r);
    r -= h1;                                                                // exact, |r| < 1
    const double h2 = trunc(r * 4294967296.0);
    r = fma(-h2, 2.3283064365386963e-10, r);                                // exact, |r| < 2^-32
    p[0] = (long long)h0; p[1] = (long long)h1; p[2] = (long long)h2; p[3] = (long long)rint(r * 18446744073709551616.0 /* 2^64 */);
}
__device__ __forceinline__ void limbs_add(LimbAcc *a, double v) {
    long long p[4];
    limbs_split(v, p);
#pragma unroll
    for (int l = 0; l < 4; ++l)
        atomicAdd(&a->w[l], (1ULL << 56) + (unsigned long long)(p[l] + (1LL << LIMB_OFFSET_LOG2)));     // result unused: RED
}
// The sum of the G contributions a limb pair received since `prev`: lane-local decoding of two words.
// ok: both count fields advanced by exactly G.  half 0: p0 2^32 + p1, half 1: p2 2^-32 + p3 2^-64; flags: 1 -Inf, 4 +Inf, 2 NaN.
__device__ __forceinline__ double limbs_decode_half(unsigned long long w0, unsigned long long w1, unsigned long long prev0, unsigned long long prev1,
                                                    int G, int half, bool &ok, unsigned &flags) {
    const unsigned long long d0 = w0 - prev0, d1 = w1 - prev1;
    ok = (int)(d0 >> 56) == G && (int)(d1 >> 56) == G;
    const long long bias = (long long)G << LIMB_OFFSET_LOG2;
    const long long s0 = (long long)(d0 & ((1ULL << 56) - 1)) - bias, s1 = (long long)(d1 & ((1ULL << 56) - 1)) - bias;
    flags = 0;
    if (half == 0) {
        if (s0 < -(LIMB_MARK >> 1)) flags |= 1u;
        if (s1 > (LIMB_MARK >> 1)) flags |= 4u;
        return fma((double)s0, 4294967296.0, (double)s1);
    }
    if (s0 > (LIMB_MARK >> 1)) flags |= 2u;
    return fma((double)s0, 2.3283064365386963e-10, (double)s1 * 5.421010862427522e-20);
}

// Worker side: fold the CTA's warp partials (last warp to arrive, fixed order) and add the CTA's sums to the chain's limb
// accumulators.  Value 0 is always delivered (an idle or commit-only pass delivers nothing else): it is the arrival signal.
__device__ __forceinline__ void cta_deliver_limbs(const Dev &d, CtaShared &sh, int c, int nc, int warp, int lane, int nworkers,
                                                  const double (&acc)[NV]) {
    // Shared-memory hand-over without a MEMBAR: the partials are written with volatile stores and the counter is bumped by
    // the same thread afterwards; shared-memory operations of a thread are performed in program order by the SM's
    // one shared-memory pipe, and the reader's loads depend on the value its own atomic returned.  (A __threadfence_block()
    // here is a MEMBAR.SC.CTA that also waits for the warp's outstanding global traffic -- the eta stores and the next
    // chain's prefetch just issued -- i.e. for a microsecond or two.)
    int last = 0;
    if (lane == 0) {
        volatile double *pw = sh.part + ((size_t)c * NWARPS + warp) * NV;
#pragma unroll
        for (int k = 0; k < NV; ++k)
            if (k < nc) pw[k] = acc[k];
        last = (atomicAdd_block(&sh.cnt[c], 1) == nworkers - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    // three lanes per value, each adding every third warp's partial, then two fixed-order adds
    static_assert(3 * NV <= 32, "three lanes per value");
    const int k0 = lane % NV, g0 = lane / NV;
    double v = 0.0;
    if (g0 < 3 && k0 < nc) {
#pragma unroll
        for (int w = 0; w < NWARPS; w += 3)
            if (w + g0 < nworkers) v += const_cast<volatile double *>(sh.part)[((size_t)c * NWARPS + w + g0) * NV + k0];
    }
    v = (v + __shfl_down_sync(0xffffffffu, v, NV)) + __shfl_down_sync(0xffffffffu, v, 2 * NV);   // valid in lanes < NV
    if (lane == 0) sh.cnt[c] = 0;
    if (lane < NV && (lane < nc || lane == 0)) limbs_add(d.lacc + (size_t)c * NV + lane, v);
    __syncwarp();
}

// Decider side: one load per lane (lanes 2k, 2k + 1: the two limb pairs of value k) fetches everything chain c's pass
// delivered; true when all `nvals` expected values are complete, out[k] then holds them (identical in every lane) and
// `prev` (the deciding warp's copy of the words, in shared memory) has moved on.  NaN / +-Inf markers become those values.
__device__ __forceinline__ bool limbs_take(const Dev &d, int c, int nvals, int lane, unsigned long long *prev /* [NV][4] */, double (&out)[NV]) {
    static_assert(2 * NV <= 32, "two lanes per value");
    const int k = lane >> 1, half = lane & 1;
    const bool mine = k < (nvals < 1 ? 1 : nvals);
    unsigned long long w0 = 0, w1 = 0;
    if (mine) {
        const unsigned long long *src = d.lacc[(size_t)c * NV + k].w + 2 * half;
        asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
    }
    bool ok = true; unsigned fl = 0; double part = 0.0;
    if (mine) part = limbs_decode_half(w0, w1, prev[k * 4 + 2 * half], prev[k * 4 + 2 * half + 1], d.G, half, ok, fl);
    if (!__all_sync(0xffffffffu, ok)) return false;
    if (mine) { prev[k * 4 + 2 * half] = w0; prev[k * 4 + 2 * half + 1] = w1; }
    const double other = __shfl_xor_sync(0xffffffffu, part, 1);
    fl |= __shfl_xor_sync(0xffffffffu, fl, 1);
    double v = half ? other + part : part + other;         // hi + lo in both lanes
    if (fl & 2u) v = NAN;
    else if ((fl & 1u) && (fl & 4u)) v = NAN;
    else if (fl & 1u) v = -INFINITY;
    else if (fl & 4u) v = INFINITY;
#pragma unroll
    for (int q = 0; q < NV; ++q) { const double t = __shfl_sync(0xffffffffu, v, 2 * q); out[q] = (q < nvals) ? t : 0.0; }
    __syncwarp();
    return true;
}

// (see jet_prepare / jet_decide below for how these are used)
constexpr int JET_R1_SO = 4;         // stepping-out tests per side in round 1
constexpr int JET_R1_PROP = 32 - 2 * JET_R1_SO;
struct JetLane { double x, l, r, pr, uA, uB; };            // this lane's point, the bracket a proposal was drawn from, log-prior at x, its uniforms
struct JetScal {
    double x0, logu, L0, R0, Jb, Kb, prior_x0, lend, rend; // lend / rend: the bracket after all JET_R1_PROP proposals were rejected
    uint64_t cursor;
    int32_t j, nAvail, openL, openR;
};
struct JetPre {                      // the deciding warp's shared-memory copy, one per chain
    double x[32], l[32], r[32], pr[32], uA[32], uB[32];
    JetScal sc;
    int32_t valid, pad;
};

// What the deciding warp of the persistent driver keeps in its CTA's shared memory between two decisions of a chain, so
// that a decision starts from shared memory instead of a chain of dependent global loads: the chain's state and control
// block as it left them, and -- fetched right AFTER a decision is published, i.e. off the critical path -- the beta /
// slice-width entries and column statistics the next decision will need.
struct DeciderCache {
    ChainState s;
    Ctl ct;
    double cst[CS_STRIDE];      // colstat row of column pref_j
    double beta_j, shat_j, beta_n, shat_n, cscale_n;
    int32_t pref_j;             // column the prefetched values belong to (-1: none)
    int32_t valid;              // s / ct hold the chain's current state
    unsigned long long prev[NV * 4];   // the chain's limb-accumulator words as of its last decided pass (limbs_take)
    JetPre pre;                 // the next update's draws and first-round points (jet_prepare), computed while its pass runs
};
__device__ __forceinline__ void decider_prefetch(const Dev &d, int c, DeciderCache *dc, int lane) {
    const int jq = dc->ct.j;
    if (jq < 0) { if (lane == 0) dc->pref_j = -1; __syncwarp(); return; }
    const int jn = (jq + 1 == d.p) ? 0 : jq + 1;
    const double *bp = d.beta + (int64_t)c * d.p, *sp = d.shat + (int64_t)c * d.p;
    double v = 0.0;
    if (lane < CS_STRIDE) v = __ldcg(d.colstat + (int64_t)jq * CS_STRIDE + lane);
    else if (lane == 12) v = __ldcg(bp + jq);
    else if (lane == 13) v = __ldcg(sp + jq);
    else if (lane == 14) v = __ldcg(bp + jn);
    else if (lane == 15) v = __ldcg(sp + jn);
    else if (lane == 16) v = __ldcg(d.colstat + (int64_t)jn * CS_STRIDE);
    if (lane < CS_STRIDE) dc->cst[lane] = v;
    else if (lane == 12) dc->beta_j = v;
    else if (lane == 13) dc->shat_j = v;
    else if (lane == 14) dc->beta_n = v;
    else if (lane == 15) dc->shat_n = v;
    else if (lane == 16) dc->cscale_n = v;
    if (lane == 0) dc->pref_j = jq;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Uniform #i (0-based within the chain's stream): replayed (R's runif record) or Philox.
__device__ __forceinline__ bool draw_uniform(const Dev &d, int c, uint64_t i, double &u) {
    if (d.replay) {
        const uint64_t r = i - d.replay_origin[c];      // the buffer of this cgg_run starts at the chain's cursor at entry
        if (r >= d.n_u) { u = 0.5; return false; }
        u = __ldcg(d.replay + (uint64_t)c * d.n_u + r);
    } else {
        u = philox_uniform(d.seed, (uint32_t)(d.chain_offset + c), i);
    }
    return true;
}

__device__ __forceinline__ int base_draws(const Dev &d) { return d.max_steps > 0 ? 3 : 2; }

// Lane-0 scalar code: fill the chain's next candidate list from its bracket state.  U[0..nU) are the
// uniforms that follow the ones already consumed by rejected shrink proposals of this update.
// Stepping-out candidates (L / R) and shrink proposals x_i = L + u_i (R - L) with the bracket update
// `if (x_i < x0) L = x_i else R = x_i` are functions of (L, R, x0, u) only, never of f (SURVEY.md fact
// 5a), so scoring several of them in one pass is exact: f only decides where the sequence stops, and
// uniforms of unused proposals are simply not consumed.
__device__ __forceinline__ void build_candidates(const Dev &d, ChainState &s, Ctl &ct, double shat, const double *U, int nU) {
    int n = 0;
    s.nL = s.nR = s.nS = 0;
    double pneed = 1.0;  // P(the next speculative shrink proposal is needed)
    if (s.phase == PH_STEPOUT) {
        // stepping-out tests: L, L - w, L - 2w, ... are known in advance too; score more than one per side when
        // expansions have been frequent (each further one is needed only if the previous was inside the slice)
        const int depth = (s.pexp > 0.6) ? 3 : ((s.pexp > 0.3) ? 2 : 1);
        if (s.openL) {
            int m = depth;
            if (d.max_steps > 0 && (double)m > s.Jb) m = (int)s.Jb;
            double v = s.L;
            for (int i = 0; i < m; ++i) { s.cand[n++] = v; s.nL++; v = __dadd_rn(v, -s.w); }
        }
        if (s.openR) {
            int m = depth;
            if (d.max_steps > 0 && (double)m > s.Kb) m = (int)s.Kb;
            double v = s.R;
            for (int i = 0; i < m; ++i) { s.cand[n++] = v; s.nR++; v = __dadd_rn(v, s.w); }
        }
        pneed = 1.0 - s.pexp;
    }
    double l = s.L, r = s.R;
    const bool prefilter = d.coarse && !s.fine_next && shat > 0.0;
    const double far = d.coarse_theta * shat;
    const bool after_undecided = s.fine_next != 0;     // the previous pass stopped at a candidate fp32 could not reject: it is probably inside
    for (int i = 0; n < d.K; ++i) {
        if (i >= nU) {
            if (i == 0 && s.phase == PH_SHRINK) s.status = CGG_E_STREAM;  // a needed draw is missing
            break;
        }
        const double x = __dadd_rn(l, __dmul_rn(U[i], __dadd_rn(r, -l)));  // L + runif(1) * (R - L)
        // speculation is a cost decision: a proposal that will go through the fp32 pre-filter costs ~1/5 of an
        // fp64 one, so it is worth scoring at a much lower probability of being needed
        const bool cheap = prefilter && fabs(x - s.x0) > far;
        if (d.tau > 0.0 && (i > 0 || s.phase == PH_STEPOUT) && pneed < (cheap ? 0.2 * d.tau : d.tau)) break;
        s.cand[n++] = x; s.nS++;
        // P(a proposal from a bracket of this width lands in the slice): shat is the bracket width at acceptance,
        // about twice the slice width, and even a bracket as tight as the slice is hit with probability < 1
        double pacc = (shat > 0.0) ? 0.5 * shat / (r - l) : 0.35;   // first sweep: no estimate yet
        pacc = pacc < 0.85 ? pacc : 0.85;
        if (after_undecided && i == 0) pacc = 0.85;
        pneed *= (1.0 - pacc);
        if (x < s.x0) l = x; else r = x;
    }
    if (s.phase == PH_STEPOUT && n == 0 && s.status == CGG_OK) s.status = CGG_E_STREAM;
    // pre-filter policy (cannot change results, only cost): a candidate further than coarse_theta slice widths
    // from x0 is almost surely outside the slice by a wide margin -> score it in fp32 and let the bound decide
    // (Tried in round 2: sending EVERY candidate through the pre-filter while a chain is far from its stationary region.  It
    // lost: from a prior draw at p = 1000 most rows sit beyond the logit clamp, the log-potential along a coordinate moves by
    // tens, not thousands, per slice width, and the rigorous fp32 bound is ~35 there -- 3.8 undecided candidates per update.)
    unsigned cm = 0;
    if (d.coarse && !s.fine_next && shat > 0.0)
        for (int k = 0; k < n; ++k)
            if (fabs(s.cand[k] - s.x0) > d.coarse_theta * shat) cm |= 1u << k;
    s.fine_next = 0;
    ct.coarse_mask = (int32_t)cm;
    s.coarse_evals += __popc(cm);
    ct.j = s.j; ct.ncand = (s.status == CGG_OK) ? n : 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) ct.delta[k] = (k < n) ? __dadd_rn(s.cand[k], -s.x0) : 0.0;
    s.cand_evals += ct.ncand;
    if (ct.ncand > 0) s.chain_passes++;
}

// Lane-0 scalar code: begin the update of coordinate s.j (qslice::slice_stepping_out up to the first
// f(L) test; R/mcmcglm.R:258-261).  f(x0) is the carried log-potential.
__device__ __forceinline__ void start_coordinate(const Dev &d, ChainState &s, double x0, const double *U, int nU) {
    s.x0 = x0;
    s.prior_rest = s.prior_sum - prior_logdens(d.prior, s.x0);
    s.sdrawn = 0; s.npass = 0;
    if (nU < base_draws(d)) { s.status = CGG_E_STREAM; return; }
    s.ylev = __dadd_rn(log(U[0]), s.fx0);                // y <- log(runif(1)) + f(x)
    s.L = __dadd_rn(s.x0, -__dmul_rn(U[1], s.w));        // L <- x - runif(1) * w
    s.R = __dadd_rn(s.L, s.w);                           // R <- L + w
    s.ref_evals += 1;                                    // qslice's f(x0)
    if (d.max_steps < 0) { s.openL = s.openR = 1; s.Jb = s.Kb = 0.0; }
    else if (d.max_steps > 0) {
        s.Jb = floor(U[2] * (double)d.max_steps);        // J <- floor(runif(1) * max)
        s.Kb = (double)d.max_steps - 1.0 - s.Jb;         // K <- max - 1 - J
        s.openL = s.Jb > 0.0; s.openR = s.Kb > 0.0;
    } else { s.openL = s.openR = 0; s.Jb = s.Kb = 0.0; }
    s.phase = (s.openL || s.openR) ? PH_STEPOUT : PH_SHRINK;
}

// The accepted value x1 of coordinate s.j (R/mcmcglm.R:262-271): beta, sample store, the eta update deferred to the
// next pass of this chain, uniform cursor, and on to the next coordinate.  `writer`: this lane does the global stores
// (the state update itself may be replicated over the lanes of the deciding warp).
__device__ __forceinline__ void accept_value(const Dev &d, int c, ChainState &s, Ctl &ct, double x1, double f, double shat_j,
                                             int shrink_draws, bool writer, double &x1_out, double &shat_out) {
    const int64_t pj = (int64_t)c * d.p + s.j;
    shat_out = (shat_j > 0.0) ? 0.75 * shat_j + 0.25 * (s.R - s.L) : (s.R - s.L);  // bracket width at acceptance
    x1_out = x1;
    if (writer) {
        d.shat[pj] = shat_out;
        d.beta[pj] = x1;                                        // R/mcmcglm.R:264
        if (d.samples) d.samples[((int64_t)c * d.n_iter + s.iter) * d.p + s.j] = x1;  // :271
    }
    ct.commit_j = s.j;                                      // eta update deferred to the next pass
    ct.commit_delta = __dadd_rn(x1, -s.x0);
    s.fx0 = f;
    s.prior_sum = s.prior_rest + prior_logdens(d.prior, x1);
    s.cursor += base_draws(d) + shrink_draws;
    s.updates++;
    s.j++;
    if (s.j == d.p) { s.j = 0; s.iter++; }
    s.phase = (s.iter >= d.iter_stop) ? PH_FLUSH : PH_START;
}

// Lane-0 scalar code: consume the log-potentials F[0..ncand) of the pass that just finished.
// Returns true when the update was accepted (s.j / s.iter advanced, s.phase = START or FLUSH);
// x1_out / shat_out then hold the accepted value and the refreshed width estimate of that coordinate.
// stop_at: index of the first pre-filtered candidate the error bound could not decide (ncand if none).  The
// sequence is consumed up to there; the chain then re-scores from that point in fp64 (fine_next).  Candidates
// the pre-filter did decide carry F = -Inf ("outside the slice").
__device__ __forceinline__ bool process_results(const Dev &d, int c, ChainState &s, Ctl &ct, const double *F, int stop_at,
                                                double shat_j, double &x1_out, double &shat_out) {
    s.npass++;
    if (ct.commit_j >= 0) { s.commit_passes++; ct.commit_j = -1; ct.commit_delta = 0.0; }
    if (s.phase == PH_FLUSH) { s.phase = PH_FINISHED; ct.ncand = 0; return false; }
    int idx = 0;
    bool expanded = false;
    if (s.phase == PH_STEPOUT) {
        // while (y < f(L)) L <- L - w   [&& J > 0 when max is finite]
        for (int i = 0; i < s.nL && s.openL; ++i) {
            if (idx + i >= stop_at) { s.fine_next = 1; s.coarse_undecided++; return false; }
            const double f = F[idx + i];
            s.ref_evals++;
            if (f != f) { s.status = CGG_E_NAN; return false; }
            if (s.ylev < f) {
                s.L = __dadd_rn(s.L, -s.w); s.stepouts++; expanded = true;
                if (d.max_steps > 0) { s.Jb -= 1.0; if (!(s.Jb > 0.0)) s.openL = 0; }
            } else s.openL = 0;
        }
        idx += s.nL;
        for (int i = 0; i < s.nR && s.openR; ++i) {
            if (idx + i >= stop_at) { s.fine_next = 1; s.coarse_undecided++; return false; }
            const double f = F[idx + i];
            s.ref_evals++;
            if (f != f) { s.status = CGG_E_NAN; return false; }
            if (s.ylev < f) {
                s.R = __dadd_rn(s.R, s.w); s.stepouts++; expanded = true;
                if (d.max_steps > 0) { s.Kb -= 1.0; if (!(s.Kb > 0.0)) s.openR = 0; }
            } else s.openR = 0;
        }
        idx += s.nR;
        s.pexp = 0.9 * s.pexp + (expanded ? 0.1 : 0.0);
        if (s.openL || s.openR) return false; // keep stepping out next pass
        s.phase = PH_SHRINK;
        if (expanded) return false;           // speculative proposals assumed the old bracket
    }
    // repeat { x1 <- L + runif(1) * (R - L); if (y < f(x1)) return x1; shrink }
    for (int i = 0; i < s.nS; ++i) {
        if (idx + i >= stop_at) { s.sdrawn += i; s.fine_next = 1; s.coarse_undecided++; return false; }
        const double f = F[idx + i], x1 = s.cand[idx + i];
        s.ref_evals++; s.shrinks++;
        if (f != f) { s.status = CGG_E_NAN; return false; }
        if (s.ylev < f) {
            accept_value(d, c, s, ct, x1, f, shat_j, s.sdrawn + i + 1, true, x1_out, shat_out);
            return true;
        }
        if (x1 < s.x0) s.L = x1; else s.R = x1;
    }
    s.sdrawn += s.nS;
    if (s.npass > 100000) s.status = CGG_E_NOTERM;
    return false;
}

// The whole slice_stepping_out update of coordinate s.j from the sums of ONE jet pass (cgg_jet.cuh).  Run by all 32
// lanes with the chain state replicated (every lane performs the same scalar updates); lanes evaluate different
// candidates: stepping-out tests L, L-w, ... (lanes 0..15) and R, R+w, ... (lanes 16..31), then 32 shrink proposals.
// The reference's test `y < f(v)` with y = log(u) + f(x0) is decided from an enclosure of f(v) (full pass: M_0 is
// the exact f(x0)) or of the difference f(v) - f(x0) (light pass: M_0 cancels, the test reads log(u) < f(v) - f(x0));
// "inside the slice" and "outside" both need the comparison to hold with the enclosure's margin.  The sequence is
// consumed while the verdicts are certain.
//   JET_ACCEPTED  the update was accepted
//   JET_EXACT     (full passes) s.phase is PH_STEPOUT / PH_SHRINK with the bracket, budgets, counters and consumed
//                 draws exactly as the reference algorithm has them at that point: the exact passes take over from there
//   JET_RETRY     (light passes) some test was not certain: the caller discards every change made here and asks for a
//                 full jet pass of the same coordinate
enum JetOutcome : int { JET_EXACT = 0, JET_ACCEPTED = 1, JET_RETRY = 2 };
// per-phase timers of a decision (CGG_PROFILE output): compiled in only with -DCGG_DECIDER_TICKS, they cost ~3 % on the
// headline workload even when profiling is off
#ifndef CGG_DECIDER_TICKS
#define CGG_TICK(slot) do { } while (0)
#else
#define CGG_TICK(slot) do { if (d.prof && lane == 0) { const long long t_ = clock64(); atomicAdd(d.prof + (slot), (unsigned long long)(t_ - tick)); tick = t_; } } while (0)
#endif

// Everything a fresh update needs that does NOT depend on the pass's sums: the uniforms, log(u), the initial bracket, and
// -- because stepping-out tests and shrink proposals are functions of (L, R, x0, u) only -- the points the first round of
// verdicts will be asked about, with their log-prior.  The deciding warp computes this right after it has published the
// previous decision, i.e. while the workers are streaming the rows, so that once the sums arrive only the enclosure
// evaluations and the comparisons remain on the chain's critical cycle.
// Lane roles of the first round: lanes 0..3 test L0 - i w, lanes 4..7 test R0 + i w, lanes 8..31 are the first 24 shrink
// proposals drawn from the bracket (L0, R0) -- valid iff stepping out does not move it, the common case.
__device__ __forceinline__ void jet_prepare(const Dev &d, int c, int lane, int j, uint64_t cursor, double x0, double w, JetLane &jl, JetScal &sc) {
    sc.x0 = x0; sc.cursor = cursor; sc.j = j;
    double uA = 0.5, uB = 0.5;
    const bool okA = draw_uniform(d, c, cursor + lane, uA);
    const bool okB = draw_uniform(d, c, cursor + 32 + lane, uB);
    const unsigned mA = __ballot_sync(0xffffffffu, okA), mB = __ballot_sync(0xffffffffu, okB);
    sc.nAvail = (mA == 0xffffffffu) ? 32 + ((mB == 0xffffffffu) ? 32 : __ffs(~mB) - 1) : __ffs(~mA) - 1;
    jl.uA = uA; jl.uB = uB;
    const double u0 = __shfl_sync(0xffffffffu, uA, 0), u1 = __shfl_sync(0xffffffffu, uA, 1), u2 = __shfl_sync(0xffffffffu, uA, 2);
    sc.logu = log(u0);                                    // y <- log(runif(1)) + f(x)
    sc.L0 = __dadd_rn(x0, -__dmul_rn(u1, w));             // L <- x - runif(1) * w
    sc.R0 = __dadd_rn(sc.L0, w);                          // R <- L + w
    sc.prior_x0 = prior_logdens(d.prior, x0);
    if (d.max_steps < 0) { sc.openL = sc.openR = 1; sc.Jb = sc.Kb = 0.0; }
    else if (d.max_steps > 0) {
        sc.Jb = floor(u2 * (double)d.max_steps);          // J <- floor(runif(1) * max)
        sc.Kb = (double)d.max_steps - 1.0 - sc.Jb;        // K <- max - 1 - J
        sc.openL = sc.Jb > 0.0; sc.openR = sc.Kb > 0.0;
    } else { sc.openL = sc.openR = 0; sc.Jb = sc.Kb = 0.0; }
    // repeat { x1 <- L + runif(1) * (R - L); ... shrink }: the proposals under the bracket (L0, R0)
    const int base = base_draws(d);
    double l = sc.L0, r = sc.R0, xi = 0.0, li = l, ri = r;
    for (int t = 0; t < JET_R1_PROP; ++t) {
        const int idx = base + t;
        const double ut = (idx < 32) ? __shfl_sync(0xffffffffu, uA, idx) : __shfl_sync(0xffffffffu, uB, idx - 32);
        const double x = __dadd_rn(l, __dmul_rn(ut, __dadd_rn(r, -l)));
        if (lane == 2 * JET_R1_SO + t) { xi = x; li = l; ri = r; }
        if (x < x0) l = x; else r = x;
    }
    sc.lend = l; sc.rend = r;
    if (lane < 2 * JET_R1_SO) {                           // stepping-out test points: L0, L0 - w, ... / R0, R0 + w, ...
        const bool left = lane < JET_R1_SO;
        const int i = left ? lane : lane - JET_R1_SO;
        double v = left ? sc.L0 : sc.R0;
        const double step = left ? -w : w;
        for (int t = 0; t < i; ++t) v = __dadd_rn(v, step);
        xi = v; li = sc.L0; ri = sc.R0;
    }
    jl.x = xi; jl.l = li; jl.r = ri;
    jl.pr = prior_logdens(d.prior, xi);
}
__device__ __forceinline__ void jet_pre_store(JetPre *p, int lane, const JetLane &jl, const JetScal &sc) {
    p->x[lane] = jl.x; p->l[lane] = jl.l; p->r[lane] = jl.r; p->pr[lane] = jl.pr; p->uA[lane] = jl.uA; p->uB[lane] = jl.uB;
    if (lane == 0) { p->sc = sc; p->valid = 1; }
    __syncwarp();
}
__device__ __forceinline__ void jet_pre_load(const JetPre *p, int lane, JetLane &jl, JetScal &sc) {
    jl.x = p->x[lane]; jl.l = p->l[lane]; jl.r = p->r[lane]; jl.pr = p->pr[lane]; jl.uA = p->uA[lane]; jl.uB = p->uB[lane];
    sc = p->sc;
}

// The verdict on one candidate v (log-prior pv) from a jet pass's sums: certainly inside the slice, certainly outside, or
// neither; fnew: the log-potential at v (full pass: enclosure midpoint; light pass: carried f(x0) + difference).  One
// routine for jet_decide and for the early publication (jet_fast_publish): the two must agree bit for bit.
struct JetJudge {
    double x0, fx0, ylev, prior_rest, logu, prior_x0, fmag, llc;
    bool light;
};
__device__ __forceinline__ void jet_verdict(const Dev &d, const double (&m)[NV], const double *cst, const JetJudge &q, double v, double pv,
                                            bool &in, bool &out, double &fnew) {
    double B;
    const double dl = jet_eval(d.family, m, cst, d.n_total, d.inv_sd, __dadd_rn(v, -q.x0), q.fmag, B, q.light, d.jet_ce);
    B = B * d.jet_bscale + 8.0 * JET_EPS * (q.fmag + fabs(dl));        // + the roundings of the sums formed below
    if (q.light) {
        const double t = dl + (pv - q.prior_x0);
        fnew = q.fx0 + t;
        in = q.logu + B < t;
        out = t + B <= q.logu;
    } else {
        // same expression order as the exact path: (ll + ll_const) + (prior_rest + prior(v)), monotone in ll
        const double ll = m[0] + dl;
        const double pr = q.prior_rest + pv;
        const double flo = ((ll - B) + q.llc) + pr, fhi = ((ll + B) + q.llc) + pr;
        fnew = (ll + q.llc) + pr;
        in = q.ylev < flo;
        out = fhi <= q.ylev;
    }
}

__device__ __forceinline__ int jet_decide(const Dev &d, int c, int lane, ChainState &s, Ctl &ct, const double (&m)[NV], bool light,
                                          double x0, double shat_j, const double *cst, const JetPre *pre,
                                          double &x1_out, double &shat_out, long long &tick) {
    s.npass++; s.jet_passes++;
    if (ct.commit_j >= 0) { s.commit_passes++; ct.commit_j = -1; ct.commit_delta = 0.0; }
    ct.coarse_mask = 0;
    const double llc = d.sharded ? 0.0 : d.ll_const;   // row-sharded: every shard's constant is already inside the exchanged M_0
    if (!light) {
        if (!(fabs(m[0]) < INFINITY)) { s.status = CGG_E_NAN; return JET_EXACT; }     // f(x0) itself is not finite
        s.fx0 = (m[0] + llc) + s.prior_sum;          // the reference's first evaluation, f(x0), at the committed eta
        // rows within reach of a link clamp (a chain still far from its stationary region): no enclosure will apply
        // until eta has moved, so do not spend jet passes on the rest of this sweep (a cost decision only)
        if (d.family != CGG_GAUSSIAN && m[9] != 0.0) s.jet_skip = 1;
    }
    const double fmag = light ? fabs(s.fx0) + 1.0 : fabs(m[0]);
    // ---- the update's draws and first-round points: prepared while the pass was running, or now
    JetLane jl; JetScal sc;
    if (pre && pre->valid && pre->sc.j == s.j && pre->sc.cursor == s.cursor && pre->sc.x0 == x0) jet_pre_load(pre, lane, jl, sc);
    else jet_prepare(d, c, lane, s.j, s.cursor, x0, s.w, jl, sc);
    CGG_TICK(14);      // uniforms and points at hand
    const int base = base_draws(d);
    // start of the update (qslice::slice_stepping_out up to the first f(L) test; R/mcmcglm.R:258-261)
    s.x0 = x0;
    s.prior_rest = s.prior_sum - sc.prior_x0;
    s.sdrawn = 0; s.npass = 0;
    if (sc.nAvail < base) { s.status = CGG_E_STREAM; return JET_EXACT; }
    s.ylev = __dadd_rn(sc.logu, s.fx0);
    s.L = sc.L0; s.R = sc.R0;
    s.ref_evals += 1;                                     // qslice's f(x0)
    s.openL = sc.openL; s.openR = sc.openR; s.Jb = sc.Jb; s.Kb = sc.Kb;
    s.phase = (s.openL || s.openR) ? PH_STEPOUT : PH_SHRINK;
    const double logu = sc.logu, prior_x0 = sc.prior_x0;
    const JetJudge judge{s.x0, s.fx0, s.ylev, s.prior_rest, logu, prior_x0, fmag, llc, light};
    auto verdict = [&](double v, double pv, bool &in, bool &out, double &fnew) { jet_verdict(d, m, cst, judge, v, pv, in, out, fnew); };
    // consume stepping-out verdicts of one side, in sequence; false: a test was not certain (the state is exact up to it)
    auto consume = [&](unsigned vin, unsigned vout, int ntests, bool left, bool &expanded) -> bool {
        for (int t = 0; t < ntests && (left ? s.openL : s.openR); ++t) {
            const bool tin = (vin >> t) & 1u, tout = (vout >> t) & 1u;
            if (!tin && !tout) return false;
            s.ref_evals++;
            if (left) {           // while (y < f(L)) L <- L - w   [&& J > 0 when max is finite]
                if (tin) { s.L = __dadd_rn(s.L, -s.w); s.stepouts++; expanded = true; if (d.max_steps > 0) { s.Jb -= 1.0; if (!(s.Jb > 0.0)) s.openL = 0; } }
                else s.openL = 0;
            } else {
                if (tin) { s.R = __dadd_rn(s.R, s.w); s.stepouts++; expanded = true; if (d.max_steps > 0) { s.Kb -= 1.0; if (!(s.Kb > 0.0)) s.openR = 0; } }
                else s.openR = 0;
            }
        }
        return true;
    };
    // ---- round 1: the first stepping-out tests of both sides AND the proposals under the initial bracket, in one go
    bool in = false, out = false; double fm = 0.0;
    {
        const bool active = (lane < JET_R1_SO) ? (s.openL != 0) : ((lane < 2 * JET_R1_SO) ? (s.openR != 0) : (base + lane - 2 * JET_R1_SO < sc.nAvail));
        if (active) verdict(jl.x, jl.pr, in, out, fm);
    }
    unsigned vin = __ballot_sync(0xffffffffu, in), vout = __ballot_sync(0xffffffffu, out);
    CGG_TICK(15);      // round 1 judged
    bool expanded = false, certain = true;
    if (s.phase == PH_STEPOUT) {
        certain = consume(vin, vout, JET_R1_SO, true, expanded);
        if (certain) certain = consume(vin >> JET_R1_SO, vout >> JET_R1_SO, JET_R1_SO, false, expanded);
        // more than JET_R1_SO expansions on a side (a slice much wider than w): further rounds of 16 tests per side
        for (int round = 0; certain && (s.openL || s.openR) && round < 64; ++round) {
            const bool left = lane < 16;
            const int i = lane & 15;
            double v = left ? s.L : s.R;
            const double step = left ? -s.w : s.w;
            for (int t = 0; t < i; ++t) v = __dadd_rn(v, step);
            bool in2 = false, out2 = false; double f2;
            if (left ? s.openL : s.openR) verdict(v, prior_logdens(d.prior, v), in2, out2, f2);
            const unsigned vi = __ballot_sync(0xffffffffu, in2), vo = __ballot_sync(0xffffffffu, out2);
            certain = consume(vi, vo, 16, true, expanded);
            if (certain) certain = consume(vi >> 16, vo >> 16, 16, false, expanded);
        }
        s.pexp = 0.9 * s.pexp + (expanded ? 0.1 : 0.0);
        if (s.openL || s.openR) {                    // a test the enclosure could not decide (or an endless expansion)
            if (light) return JET_RETRY;
            s.jet_fallbacks++;
            return JET_EXACT;
        }
        s.phase = PH_SHRINK;
    }
    // ---- the shrink proposals: round 1's if the bracket did not move, else drawn again from the expanded bracket
    int nprop = JET_R1_PROP, poff = 2 * JET_R1_SO;
    double lend = sc.lend, rend = sc.rend, xi = jl.x, li = jl.l, ri = jl.r;
    if (expanded) {
        nprop = 32; poff = 0;
        const int sidx = base + lane;
        const double usA = __shfl_sync(0xffffffffu, jl.uA, sidx & 31), usB = __shfl_sync(0xffffffffu, jl.uB, sidx & 31);
        const double us = (sidx < 32) ? usA : usB;
        double l = s.L, r = s.R;
        for (int t = 0; t < 32; ++t) {
            const double ut = __shfl_sync(0xffffffffu, us, t);
            const double x = __dadd_rn(l, __dmul_rn(ut, __dadd_rn(r, -l)));
            if (lane == t) { xi = x; li = l; ri = r; }
            if (x < s.x0) l = x; else r = x;
        }
        lend = l; rend = r;
        in = false; out = false; fm = 0.0;
        if (sidx < sc.nAvail) verdict(xi, prior_logdens(d.prior, xi), in, out, fm);
        vin = __ballot_sync(0xffffffffu, in); vout = __ballot_sync(0xffffffffu, out);
    }
    CGG_TICK(16);      // stepping out decided, proposals judged
    const unsigned pmask = (nprop == 32) ? 0xffffffffu : ((1u << nprop) - 1u);
    const unsigned pout = (vout >> poff) & pmask, pin = (vin >> poff) & pmask;
    const int k = (pout == pmask) ? nprop : __ffs(~pout) - 1;      // first proposal that is not certainly rejected
    if (k == nprop) {
        if (light) return JET_RETRY;
        s.shrinks += nprop; s.ref_evals += nprop; s.sdrawn = nprop; s.L = lend; s.R = rend;
        s.jet_fallbacks++;
        return JET_EXACT;
    }
    const bool kin = (pin >> k) & 1u;
    s.L = __shfl_sync(0xffffffffu, li, k + poff); s.R = __shfl_sync(0xffffffffu, ri, k + poff);   // the bracket proposal k was drawn from
    s.shrinks += k; s.ref_evals += k;
    if (!kin) {                 // undecided (or its draw is not available): the exact passes continue from proposal k
        if (light) return JET_RETRY;
        s.sdrawn = k;
        s.jet_fallbacks++;
        return JET_EXACT;
    }
    s.shrinks++; s.ref_evals++;
    const double x1 = __shfl_sync(0xffffffffu, xi, k + poff), f1 = __shfl_sync(0xffffffffu, fm, k + poff);
    accept_value(d, c, s, ct, x1, f1, shat_j, k + 1, lane == 0, x1_out, shat_out);
    CGG_TICK(17);      // value accepted
    return JET_ACCEPTED;
}

// THE PLAIN UPDATE, decided and booked without the general machinery.  In the steady state an update is: no stepping-out
// expansion, and the first shrink proposal that is not certainly rejected is certainly accepted -- all of it visible in the
// ROUND-1 verdicts, whose points were prepared while the pass was streaming.  For that case (and only when nothing else
// is special: not the last column of an iteration, no clamp-risk rows, the whole replay window available) the deciding
// warp works straight on its shared-memory cache:
//   1. judges round 1 (jet_verdict: the same routine, hence the same verdicts, as jet_decide);
//   2. persistent driver: writes the four control words the workers need -- next column, pending column, its delta, the
//      column scale -- and releases the chain's version, ~2 us after the sums arrived (`ver` != ~0);
//   3. books the update: exactly the state changes, counters and global stores (beta, slice-width estimate, sample) that
//      decide_chain / jet_decide / accept_value make for this case, field by field, without copying the 300-byte chain state
//      into registers and back (the general path spends ~20k cycles per decision, most of it such traffic).
// Returns false -- nothing changed -- whenever the update is not of the plain kind; the general path then runs as ever.
// CGG_EARLY=0 turns this off; every parity test compares chains with it (the default) against the oracle, and the
// all-exact engine, the counters included.
__device__ __forceinline__ bool jet_fast_update(const Dev &d, int c, int lane, DeciderCache *dc, const double (&m)[NV], bool light,
                                                unsigned long long ver) {
    const JetPre *pre = &dc->pre;
    const int j = dc->s.j;
    if (!(dc->pref_j == j && dc->ct.j == j && pre->valid && pre->sc.j == j && pre->sc.cursor == dc->s.cursor && pre->sc.x0 == dc->beta_j)) return false;
    if (dc->s.status != CGG_OK || dc->s.jet_skip || !d.jet) return false;
    if ((int64_t)j + 1 >= d.p) return false;                 // the iteration ends with this update (flush, prior re-sum, ...)
    if (pre->sc.nAvail < 64) return false;                   // a replayed stream that is running out
    JetJudge q;
    q.light = light; q.x0 = dc->beta_j; q.logu = pre->sc.logu; q.prior_x0 = pre->sc.prior_x0;
    q.llc = d.sharded ? 0.0 : d.ll_const;
    q.prior_rest = dc->s.prior_sum - pre->sc.prior_x0;       // exactly jet_decide's expressions
    if (light) { q.fx0 = dc->s.fx0; q.fmag = fabs(q.fx0) + 1.0; }
    else {
        if (!(fabs(m[0]) < INFINITY)) return false;
        if (d.family != CGG_GAUSSIAN && m[9] != 0.0) return false;
        q.fx0 = (m[0] + q.llc) + dc->s.prior_sum;
        q.fmag = fabs(m[0]);
    }
    q.ylev = __dadd_rn(pre->sc.logu, q.fx0);
    const bool openL = pre->sc.openL != 0, openR = pre->sc.openR != 0;
    const double x = pre->x[lane];
    bool in = false, out = false; double fm = 0.0;
    if ((lane < JET_R1_SO) ? openL : ((lane < 2 * JET_R1_SO) ? openR : true)) jet_verdict(d, m, dc->cst, q, x, pre->pr[lane], in, out, fm);
    const unsigned vin = __ballot_sync(0xffffffffu, in), vout = __ballot_sync(0xffffffffu, out);
    if (openL && !(vout & 1u)) return false;                                 // f(L0) is not certainly below the level: expansion, or undecided
    if (openR && !((vout >> JET_R1_SO) & 1u)) return false;
    const unsigned pmask = (1u << JET_R1_PROP) - 1u;
    const unsigned pout = (vout >> (2 * JET_R1_SO)) & pmask, pin = (vin >> (2 * JET_R1_SO)) & pmask;
    if (pout == pmask) return false;
    const int k = __ffs(~pout) - 1;
    if (!((pin >> k) & 1u)) return false;
    const int src = k + 2 * JET_R1_SO;
    const double x1 = __shfl_sync(0xffffffffu, x, src), f1 = __shfl_sync(0xffffffffu, fm, src);
    const double lk = __shfl_sync(0xffffffffu, pre->l[lane], src), rk = __shfl_sync(0xffffffffu, pre->r[lane], src);     // the bracket proposal k was drawn from
    if (lane == 0) {
        const bool full_next = !d.jet_light || d.family != CGG_BINOMIAL;
        const int mask_next = (int)(JET_BIT | (full_next ? JET_FULL : 0u));
        const double cdelta = __dadd_rn(x1, -q.x0);
        if (ver != ~0ULL) {         // persistent driver: the workers may go on
            Ctl *g = d.ctl + c;
            *reinterpret_cast<int4 *>(g) = make_int4(j + 1, 0, j, mask_next);      // j, ncand, commit_j, coarse_mask
            g->commit_delta = cdelta;
            g->cscale = dc->cscale_n;
            st_release_u64(&d.sync[c].version, ver + 1ULL);
        }
        // ---- the book-keeping of decide_chain / jet_decide / accept_value for this case
        ChainState &S = dc->s;
        Ctl &T = dc->ct;
        S.passes++; S.jet_passes++;
        if (T.commit_j >= 0) S.commit_passes++;
        S.x0 = q.x0; S.prior_rest = q.prior_rest; S.sdrawn = 0; S.npass = 0;
        S.ylev = q.ylev;
        S.ref_evals += 1u + (openL ? 1u : 0u) + (openR ? 1u : 0u) + (unsigned)(k + 1);     // f(x0), the stepping-out tests, the proposals
        S.shrinks += (unsigned)(k + 1);
        S.openL = 0; S.openR = 0; S.Jb = pre->sc.Jb; S.Kb = pre->sc.Kb;
        if (openL || openR) S.pexp = 0.9 * S.pexp;
        S.L = lk; S.R = rk;
        const double shat_j = dc->shat_j;
        const double shat_out = (shat_j > 0.0) ? 0.75 * shat_j + 0.25 * (rk - lk) : (rk - lk);
        const int64_t pj = (int64_t)c * d.p + j;
        d.shat[pj] = shat_out;
        d.beta[pj] = x1;                                        // R/mcmcglm.R:264
        if (d.samples) d.samples[((int64_t)c * d.n_iter + S.iter) * d.p + j] = x1;  // :271
        S.fx0 = f1;
        S.prior_sum = q.prior_rest + prior_logdens(d.prior, x1);
        S.cursor += base_draws(d) + k + 1;
        S.updates++;
        S.j = j + 1;
        S.phase = PH_JET; S.chain_passes++;
        T.j = j + 1; T.ncand = 0; T.commit_j = j; T.coarse_mask = mask_next; T.commit_delta = cdelta; T.cscale = dc->cscale_n;
    }
    __syncwarp();
    return true;
}

// Row-sharded persistent driver: the ranks' totals of a pass, exchanged through peer-mapped mailboxes and added in RANK
// ORDER, so that every rank decides from bit-identical sums.  Each value travels as two self-validating 64-bit words
// {stamp32 : half32} (cf. LimbAcc: a 64-bit element of a vector access cannot be torn), written straight into every
// peer's memory with system-scope stores -- NVLink on an NVSwitch box -- and polled for locally: no collective call, no
// host, no extra launch; the pass -> exchange -> decision loop never leaves the persistent kernel.  stamp = number of
// the pass since the handle was created (never reset: nothing has to be cleared between runs).
struct __align__(16) MboxEntry { unsigned long long w0, w1; };
__device__ __forceinline__ void mbox_store(MboxEntry *dst, double v, unsigned stamp) {
    const unsigned long long s = (unsigned long long)stamp << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(s | (unsigned)__double2loint(v)), "l"(s | (unsigned)__double2hiint(v)) : "memory");
}
__device__ __forceinline__ bool mbox_load(const MboxEntry *src, unsigned stamp, double &v) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
    v = __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
    return (unsigned)(w0 >> 32) == stamp && (unsigned)(w1 >> 32) == stamp;
}
__device__ __forceinline__ void mbox_load_raw(const MboxEntry *src, unsigned long long &w0, unsigned long long &w1) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
}
__device__ __forceinline__ bool mbox_decode(unsigned long long w0, unsigned long long w1, unsigned stamp, double &v) {
    v = __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
    return (unsigned)(w0 >> 32) == stamp && (unsigned)(w1 >> 32) == stamp;
}
// vals[0..nvals): in: this rank's sums (identical in every lane); out: the totals over the ranks.  false: timed out.
__device__ __forceinline__ bool mbox_exchange(const Dev &d, int c, int nvals, unsigned stamp, int lane, double (&vals)[NV]) {
    const int nv = nvals < 1 ? 1 : nvals;
    // my values to everybody (lane k sends value k to every rank, my own mailbox included)
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) mine = (lane == k) ? vals[k] : mine;
    if (lane < nv)
        for (int r = 0; r < d.world; ++r)
            mbox_store(reinterpret_cast<MboxEntry *>(d.mbox[r]) + ((size_t)c * d.world + d.rank) * NV + lane, mine, stamp);
    // everybody's values from my mailbox: the world * nv entries are spread over the lanes (up to three per lane, requested
    // back to back: one round trip per look), then lane k adds value k of the ranks in rank order
    const MboxEntry *my = reinterpret_cast<const MboxEntry *>(d.mbox[d.rank]) + (size_t)c * d.world * NV;
    const int total = d.world * nv;                 // <= 8 * 10
    const unsigned long long t0 = globaltimer_ns();
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    for (unsigned spins = 0;; ++spins) {
        const int i0 = lane, i1 = lane + 32, i2 = lane + 64;
        // the three requests first, the checks after (a check right behind its load would serialise the round trips)
        const unsigned long long sw = ((unsigned long long)stamp << 32);
        unsigned long long a0 = sw, a1 = sw, b0 = sw, b1 = sw, c0 = sw, c1 = sw;
        if (i0 < total) mbox_load_raw(my + (size_t)(i0 / nv) * NV + (i0 % nv), a0, a1);
        if (i1 < total) mbox_load_raw(my + (size_t)(i1 / nv) * NV + (i1 % nv), b0, b1);
        if (i2 < total) mbox_load_raw(my + (size_t)(i2 / nv) * NV + (i2 % nv), c0, c1);
        const bool ok = mbox_decode(a0, a1, stamp, e0) & mbox_decode(b0, b1, stamp, e1) & mbox_decode(c0, c1, stamp, e2);
        if (__all_sync(0xffffffffu, ok)) break;
        if ((spins & 1023u) == 1023u && globaltimer_ns() - t0 > 4000000000ULL) return false;      // a peer never showed up
    }
    double tot = 0.0;
    for (int r = 0; r < d.world; ++r) {
        const int idx = r * nv + (lane < nv ? lane : 0);
        const double a0 = __shfl_sync(0xffffffffu, e0, idx & 31), a1 = __shfl_sync(0xffffffffu, e1, idx & 31), a2 = __shfl_sync(0xffffffffu, e2, idx & 31);
        tot += (idx < 32) ? a0 : ((idx < 64) ? a1 : a2);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) vals[k] = __shfl_sync(0xffffffffu, tot, k);
    return true;
}

// Persistent driver, right after a decision was published and the next column's beta / statistics were prefetched: if the
// pass now in flight is a jet pass that starts a fresh update, prepare that update (jet_prepare) while the rows stream.
__device__ __forceinline__ void decider_prephase(const Dev &d, int c, DeciderCache *dc, int lane) {
    const bool want = dc->valid && dc->s.status == CGG_OK && dc->s.phase == PH_JET && dc->ct.j >= 0 && dc->pref_j == dc->ct.j;
    if (!want) { if (lane == 0) dc->pre.valid = 0; __syncwarp(); return; }
    if (dc->pre.valid && dc->pre.sc.j == dc->s.j && dc->pre.sc.cursor == dc->s.cursor && dc->pre.sc.x0 == dc->beta_j) return;   // (a retry of the same update)
    JetLane jl; JetScal sc;
    jet_prepare(d, c, lane, dc->s.j, dc->s.cursor, dc->beta_j, dc->s.w, jl, sc);
    jet_pre_store(&dc->pre, lane, jl, sc);
}

// One warp decides one chain after every worker's contribution to the pass is visible.
// j_hint: the column of the pass just finished if the caller already knows it (>= 0), else -1.
// from_xbuf: log-likelihood sums come from d.xbuf (already reduced across ranks, row-sharded mode).
// Returns true when the chain is finished (or failed) after this decision.
// The parameter block is taken by pointer -- the kernels pass the address of their __grid_constant__
// parameter -- so this cold, register-hungry routine stays out of line and off the hot loop's registers.
enum SumSource : int { SRC_ACC = 0, SRC_XBUF = 1, SRC_SLOTS = 2 };
// exchanged sum #idx of the row-sharded mode: already reduced in d.xbuf (host hook), or the ranks' parts added in rank order
__device__ __forceinline__ double xbuf_value(const Dev &d, int idx) {
    if (!d.gathered) return __ldcg(d.xbuf + idx);
    double v = 0.0;
    for (int r = 0; r < d.world; ++r) v += __ldcg(d.gathered + (size_t)r * d.C * NV + idx);
    return v;
}
// vals_in (with SRC_SLOTS): the pass's sums are handed over by the caller (cluster driver) instead of read from the limbs.
enum DecideOutcome : int { DEC_CONTINUE = 0, DEC_FINISHED = 1, DEC_NOT_READY = 2, DEC_ABORT = 3, DEC_PUBLISHED = 4 /* continue; the version is already released */ };
__device__ __noinline__ int decide_chain(const Dev *dp, int c, int lane, int j_hint, int src, DeciderCache *dc = nullptr,
                                         const double *vals_in = nullptr, unsigned pass_no = 0, unsigned long long ver = ~0ULL) {
    const bool from_xbuf = src == SRC_XBUF;
    const Dev &d = *dp;
    // ---- control block, state, beta/shat of j and j+1: from the deciding warp's shared-memory cache if it has them,
    // else one batched round of global loads
    const bool cached = dc && dc->valid;
    if (cached) j_hint = dc->ct.j;
    else if (j_hint < 0) j_hint = __ldcg(&d.ctl[c].j);
    const int jq = j_hint < 0 ? 0 : j_hint;              // column of the pass that just finished
    const int jn = (jq + 1 == d.p) ? 0 : jq + 1;         // the column an acceptance moves on to
    const double *bp = d.beta + (int64_t)c * d.p, *sp = d.shat + (int64_t)c * d.p;
    const bool pref = dc && dc->pref_j == jq;
    const double beta_j = pref ? dc->beta_j : __ldcg(bp + jq), beta_n = pref ? dc->beta_n : __ldcg(bp + jn);
    const double shat_j = pref ? dc->shat_j : __ldcg(sp + jq), shat_n = pref ? dc->shat_n : __ldcg(sp + jn);
    const double *cst_j = pref ? dc->cst : d.colstat + (int64_t)jq * CS_STRIDE;          // statistics of column jq
    const double cscale_n = pref ? dc->cscale_n : __ldcg(d.colstat + (int64_t)jn * CS_STRIDE);
    // (the 400 bytes of control block and chain state are only copied once the plain update -- jet_fast_update, which works
    // on the cache itself -- has been ruled out: until then four scalars of them are all that is looked at)
    const Ctl *ctp = cached ? &dc->ct : &d.ctl[c];
    const ChainState *stp = cached ? &dc->s : &d.cs[c];
    const int ph0 = stp->phase, st0 = stp->status, nc0 = ctp->ncand;
    const unsigned cm0 = (unsigned)ctp->coarse_mask;
#ifdef CGG_DECIDER_TICKS
    long long tick = d.prof ? clock64() : 0;
#else
    long long tick = 0;
#endif
    if (ph0 == PH_FINISHED || st0 != CGG_OK) return DEC_FINISHED;
    const bool jetpass = (cm0 & JET_BIT) != 0u && ph0 == PH_JET;
    const int nc = jetpass ? 0 : nc0;
    const unsigned cmask = jetpass ? 0u : cm0;
    double jm[NV];        // the pass's sums, identical in every lane (slots source: all of them; else only for a jet pass)
    if (src == SRC_SLOTS) {
        const int nvals = jetpass ? jet_nvals(d.family, (cm0 & JET_FULL) == 0u) : (cmask ? nc + 2 : nc);
        if (vals_in) {
#pragma unroll
            for (int k = 0; k < NV; ++k) jm[k] = (k < nvals) ? vals_in[k] : 0.0;
        } else if (!limbs_take(d, c, nvals, lane, dc->prev, jm)) return DEC_NOT_READY;  // some CTA's sums are still on their way: nothing was changed
        if (d.sharded && d.mbox[0]) {
            // row-sharded: this shard's additive constant joins its sums (M_0 of a jet pass that delivers it, every candidate
            // sum of an exact pass), then the ranks' sums are exchanged and added in rank order
            const bool has_m0 = jetpass && (d.family != CGG_BINOMIAL || (cm0 & JET_FULL) != 0u);
#pragma unroll
            for (int k = 0; k < NV; ++k) if (jetpass ? (k == 0 && has_m0) : (k < nc)) jm[k] += d.ll_const;
            if (!mbox_exchange(d, c, nvals, d.mbox_stamp0 + pass_no + 1u, lane, jm)) return DEC_ABORT;
        }
    } else if (jetpass) {
        const double mv = (lane < NV) ? (from_xbuf ? xbuf_value(d, c * NV + lane) : acc_take(d.acc + c * NV + lane)) : 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) jm[k] = __shfl_sync(0xffffffffu, mv, k);
    }
    if (jetpass && d.family == CGG_BINOMIAL && (cm0 & JET_FULL) == 0u) jet_light_unpack(jm);
    // ---- the plain update: judged, published and booked straight from the deciding warp's cache (jet_fast_update)
    if (jetpass && src == SRC_SLOTS && cached && d.early &&
        jet_fast_update(d, c, lane, dc, jm, (cm0 & JET_FULL) == 0u, vals_in ? ~0ULL : ver))
        return (!vals_in && ver != ~0ULL) ? DEC_PUBLISHED : DEC_CONTINUE;
    Ctl ct = *ctp;
    ChainState s = *stp;
    if (s.x0 != s.x0) tick = 0;   // (keeps the state loads above the first timestamp)
    CGG_TICK(12);      // state loaded, sums read
    // lane k: total log-likelihood of candidate k + its prior term
    double f = 0.0;
    unsigned int aflags = 0;
    auto pick = [&](int idx) { double v = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) v = (k == idx) ? jm[k] : v;
        return v; };
    if (lane < nc) {
        double ll;
        if (src == SRC_SLOTS) {
            ll = pick(lane) + (d.sharded ? 0.0 : d.ll_const);       // (row-sharded: the constants travelled with the sums)
            if (cmask) { aflags = __ldcg(&d.acc[c * NV + lane].flags); d.acc[c * NV + lane].flags = 0u; }   // clamp-proximity flag of the pre-filter
        } else ll = from_xbuf ? xbuf_value(d, c * NV + lane) : acc_take(d.acc + c * NV + lane, &aflags) + d.ll_const;
        f = ll + (s.prior_rest + prior_logdens(d.prior, s.cand[lane]));
    }
    int stop_at = nc;
    if (cmask) {
        // pre-filtered candidates: |f32 sum - exact| <= B.  The candidate is outside the slice for certain iff
        // f + B < ylev; anything else (including a row on the clamp discontinuity, or NaN) stays undecided.
        double bsum = 0.0;
        if (src != SRC_SLOTS && (lane == nc || lane == nc + 1)) bsum = acc_take(d.acc + c * NV + lane);
        const double sE = (src == SRC_SLOTS) ? pick(nc) : __shfl_sync(0xffffffffu, bsum, nc);
        const double sX = (src == SRC_SLOTS) ? pick(nc + 1) : __shfl_sync(0xffffffffu, bsum, nc + 1);
        bool undecided = false;
        if (lane < nc && ((cmask >> lane) & 1u)) {
            const double B = 1.01 * ((2.384185791015625e-07 + (double)kCoarseKappa) * (sE + fabs(s.cand[lane] - s.x0) * sX)
                                     + (double)kCoarseKappa * (double)d.n);
            // certainly outside the slice: y >= f for any exact value within B.  A STEPPING-OUT test only needs the verdict
            // (`while (y < f(L)) L <- L - w` never uses f itself), so it may also be certainly INSIDE; a shrink proposal that
            // is inside is the accepted one and needs its exact f (the next update's f(x0)): it stays undecided.
            const bool clean = !(aflags & 8u);
            const bool out = clean && (f + B < s.ylev);
            const bool in_ = clean && (s.ylev < f - B) && s.phase == PH_STEPOUT && lane < s.nL + s.nR;
            undecided = !(out || in_);
            if (out) f = -INFINITY; else if (in_) f = INFINITY;
        }
        const unsigned um = __ballot_sync(0xffffffffu, undecided);
        if (um) stop_at = __ffs(um) - 1;
    }
    double F[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) F[k] = __shfl_sync(0xffffffffu, f, k);
    s.passes++;
    double x0 = beta_j, shat = shat_j;     // values of the coordinate that is sampled next
    bool retry_full = false;
    if (jetpass) {
        double x1 = 0.0, sh1 = 0.0;
        const bool light = ((unsigned)ct.coarse_mask & JET_FULL) == 0u;
        const int oc = jet_decide(d, c, lane, s, ct, jm, light, beta_j, shat_j, cst_j, dc ? &dc->pre : nullptr, x1, sh1, tick);
        if (oc == JET_ACCEPTED) {
            if (jn != jq) { x0 = beta_n; shat = shat_n; } else { x0 = x1; shat = sh1; }
        } else if (oc == JET_RETRY) {
            // a light pass left a test undecided: forget everything it changed, keep only the facts of the pass itself
            // (the pending eta update has been applied), and ask for a full jet pass of the same coordinate
            s = cached ? dc->s : d.cs[c]; ct = cached ? dc->ct : d.ctl[c];
            s.npass++; s.jet_passes++; s.jet_retries++;
            if (ct.commit_j >= 0) { s.commit_passes++; ct.commit_j = -1; ct.commit_delta = 0.0; }
            retry_full = true;
        }
    } else if (lane == 0 && s.phase != PH_START) {
        double x1 = 0.0, sh1 = 0.0;
        if (process_results(d, c, s, ct, F, stop_at, shat_j, x1, sh1)) {
            if (jn != jq) { x0 = beta_n; shat = shat_n; } else { x0 = x1; shat = sh1; }   // p == 1: same column again
        }
    }
    const int phase = __shfl_sync(0xffffffffu, s.phase, 0);
    const int status = __shfl_sync(0xffffffffu, s.status, 0);
    const int j = __shfl_sync(0xffffffffu, s.j, 0);
    __syncwarp();
    if (status == CGG_OK && phase == PH_START && j == 0) {
        // once per sweep: re-sum the prior over all p coordinates (quirk Q5) so the running value cannot drift
        double v = 0.0;
        for (int64_t l = lane; l < d.p; l += 32) v += prior_logdens(d.prior, __ldcg(bp + l));
        v = warp_sum(v);
        s.prior_sum = v;
    }
    const int jskip = __shfl_sync(0xffffffffu, (j == 0 && phase == PH_START) ? 0 : s.jet_skip, 0);   // a new sweep tries again
    if (lane == 0) s.jet_skip = jskip;
    if (status == CGG_OK && d.jet && ((phase == PH_START && !jskip) || retry_full)) {
        // jet mode: the next pass of this chain applies the pending eta update and delivers the derivative moments along
        // the new column (and the exact f(x0) if it is a full pass); the whole update is then decided from them
        if (lane == 0) {
            const bool full = retry_full || !d.jet_light || d.family != CGG_BINOMIAL;
            ct.j = s.j; ct.ncand = 0; ct.coarse_mask = (int32_t)(JET_BIT | (full ? JET_FULL : 0u));
            ct.cscale = (s.j == jn) ? cscale_n : ((s.j == jq) ? cst_j[0] : __ldcg(d.colstat + (int64_t)s.j * CS_STRIDE));
            s.phase = PH_JET; s.chain_passes++;
        }
    } else if (status == CGG_OK && (phase == PH_START || phase == PH_SHRINK || phase == PH_STEPOUT)) {
        // uniforms the next decision steps can need, fetched by the lanes in parallel
        const uint64_t cursor = __shfl_sync(0xffffffffu, (unsigned long long)s.cursor, 0);
        const int sdrawn = __shfl_sync(0xffffffffu, s.sdrawn, 0);
        const uint64_t ustart = (phase == PH_START) ? cursor : cursor + base_draws(d) + sdrawn;
        double u = 0.5;
        const bool ok = (lane < NU) ? draw_uniform(d, c, ustart + lane, u) : true;
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        int nU = __ffs(~okmask) - 1;   // first lane whose draw was unavailable
        if (nU < 0 || nU > NU) nU = NU;
        double U[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) U[i] = __shfl_sync(0xffffffffu, u, i);
        if (lane == 0) {
            int off = 0;
            if (phase == PH_START) { start_coordinate(d, s, x0, U, nU); off = base_draws(d); }
            if (s.status == CGG_OK) build_candidates(d, s, ct, shat, U + off, nU - off > 0 ? nU - off : 0);
        }
    } else if (lane == 0) {
        ct.ncand = 0; ct.coarse_mask = 0;
    }
    bool fin = false;
    if (lane == 0) {
        if (s.status != CGG_OK) { ct.ncand = 0; ct.commit_j = -1; }
        fin = (s.phase == PH_FINISHED) || (s.status != CGG_OK);
        if (fin) ct.j = -1;
        // the deciding warp of the persistent driver keeps the state in shared memory; global memory gets it when the
        // chain stops (the host reads it after the kernel) -- every decision otherwise
        if (!dc || fin) d.cs[c] = s;
        d.ctl[c] = ct;
        if (dc) { dc->s = s; dc->ct = ct; dc->valid = 1; }
    }
    fin = __shfl_sync(0xffffffffu, (int)fin, 0);
    CGG_TICK(18);      // next pass set up, state stored
    // every lane that cleared accumulators (or their flags) must have that visible before the version is released; with
    // the slot source and no pre-filter only lane 0 wrote, and its release store orders its own writes
    if (!(src == SRC_SLOTS && cmask == 0u)) fence_gpu();
    __syncwarp();
    CGG_TICK(19);      // fence
    return fin ? DEC_FINISHED : DEC_CONTINUE;
}

}  // namespace cgg
