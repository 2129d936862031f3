// cggibbs.cu -- kernels and C ABI of libcggibbs.so (see include/cggibbs.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "cgg_device.cuh"
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved with dlopen at run time

using namespace cgg;

// ============================================================================================
// Kernels
// ============================================================================================

// K3 (persistent driver).  Launched cooperatively, one 8-warp CTA per SM, so that every warp of the grid is resident:
// d.G worker CTAs and, behind them, ONE decider CTA.
//   worker warps  own a fixed set of 64-row tiles for the whole run and walk the chains round-robin.  A chain's pass #r
//                 may start as soon as its decision #r is published (version >= r).  Kinds of walk: a single chain
//                 (worker_pass: jet or exact pass), a pair of chains at the same coordinate through the general loop
//                 (warp_pass_jet2: any operand source), a pair through the lean loop (warp_pass_group<.., 2>: the steady
//                 state, everything but eta from the X-column cache), and -- builds with -DCGG_GROUP_PASSES -- four chains.
//                 A warp's sums go to its CTA without blocking; the last warp of the CTA to arrive folds them in warp order
//                 and adds the CTA's sums to the chain's limb accumulators (cta_deliver_limbs): no fence, no flag.
//   decider warps (warp w of the decider CTA decides chains w, w + 8, ...) watch their chains' accumulators and the moment
//                 a pass is complete either take the plain-update path (jet_fast_update: judge round 1, publish, book) or
//                 run the general state machine (decide_chain), publish the next version, and then prepare the chain's next
//                 decision while its pass streams (decider_prefetch, decider_prephase).
// There is no grid-wide barrier and no worker ever runs a decision: decisions of different chains proceed concurrently,
// and with >= 3 pairs of chains per device their latency hides behind the streaming of the other pairs.
constexpr unsigned long long VERSION_FINISHED = 1ULL << 62;

__device__ __forceinline__ bool wait_timed_out(const Dev &d, unsigned long long t0, int lane) {
    if (__ldcg(&d.hdr->abort)) return true;
    if (globaltimer_ns() - t0 > 4000000000ULL) {  // a peer never showed up: bail out, do not hang the GPU
        if (lane == 0) { d.hdr->abort = 1; fence_gpu(); }
        return true;
    }
    return false;
}

// The deciding warp of chains first_chain, first_chain + stride, ...: as soon as every worker CTA's slot of a chain
// carries the stamp of the pass in flight, decide and publish; then fetch what the chain's next decision will need.
__device__ __noinline__ void decider_loop(const Dev *dp, int first_chain, int stride, int lane, DeciderCache *cache) {
    const Dev &d = *dp;
    for (int c = first_chain; c < d.C; c += stride) {
        if (lane == 0) { cache[c].valid = 0; cache[c].pref_j = -1; cache[c].pre.valid = 0; }
        for (int i = lane; i < NV * 4; i += 32) cache[c].prev[i] = 0ULL;      // the accumulators are zeroed before every launch
    }
    __syncwarp();
    unsigned long long t0 = globaltimer_ns();
    unsigned idle = 0;
    for (;;) {
        bool all_fin = true, progressed = false;
        for (int c = first_chain; c < d.C; c += stride) {
            unsigned long long v = 0;
            if (lane == 0) v = __ldcg(&d.sync[c].version);
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v >= VERSION_FINISHED) continue;
            all_fin = false;
#ifdef CGG_PROFILE_BUILD
            const bool dprof = d.prof != nullptr;
#else
            constexpr bool dprof = false;
#endif
            const unsigned long long tg0 = dprof ? globaltimer_ns() : 0;
            const long long td0 = dprof ? clock64() : 0;
            const int oc = decide_chain(dp, c, lane, -1, SRC_SLOTS, cache + c, nullptr, (unsigned)v, v);  // reads the limb accumulators itself: one load per lane
            if (oc == DEC_NOT_READY) continue;                                      // pass #v still has CTAs streaming
            if (oc == DEC_ABORT) { if (lane == 0) { d.hdr->abort = 1; fence_gpu(); } return; }   // a peer rank never delivered
            const bool fin = oc == DEC_FINISHED;
            if (lane == 0 && oc != DEC_PUBLISHED) st_release_u64(&d.sync[c].version, fin ? VERSION_FINISHED : v + 1);
            if (!fin) {                                              // after publishing: off the chain's critical path
                decider_prefetch(d, c, cache + c, lane);
                decider_prephase(d, c, cache + c, lane);
            }
            if (dprof && lane == 0) {
                atomicAdd(d.prof + 7, (unsigned long long)(clock64() - td0)); atomicAdd(d.prof + 8, 1ULL);
                if (v < 128) { d.prof[32 + 4096 + (c * 128 + v) * 4 + 0] = tg0; d.prof[32 + 4096 + (c * 128 + v) * 4 + 1] = globaltimer_ns(); }
            }
            progressed = true;
        }
        if (all_fin) break;
        if (progressed) { idle = 0; t0 = globaltimer_ns(); }
        else {
            __nanosleep(40);
            if ((++idle & 255u) == 0 && wait_timed_out(d, t0, lane)) {
                break;
            }
        }
    }
}

template <int FAMILY>
__global__ void __launch_bounds__(THREADS, 1) sweep_persistent_kernel(const __grid_constant__ Dev d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double2 s_l1p[MATH_TAB_N];
    CtaShared sh(smem_raw, d.C);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (__ldcg(&d.hdr->abort)) return;        // an earlier launch of this run gave up
    load_l1p_table(s_l1p);
    for (int i = threadIdx.x; i < d.C; i += THREADS) { sh.ver[i] = 0ULL; sh.cnt[i] = 0; sh.lock[i] = 0; }
    for (int i = threadIdx.x; i < d.C * CTL_WORDS; i += THREADS) sh.ctl[i] = __ldcg(reinterpret_cast<const double *>(d.ctl) + i);
    __syncthreads();
    // The deciders have a CTA (an SM) of their own, the one after the d.G worker CTAs: warp w decides chains w, w + NWARPS, ...
    // Next to fifteen workers that saturate the fp64 pipe a decision took ~25 us, and the decision sits on every
    // chain's critical cycle (pass -> decision -> next pass).
    if ((int)blockIdx.x == d.G) {
        if (warp < d.C) decider_loop(&d, warp, NWARPS, lane, reinterpret_cast<DeciderCache *>(smem_raw));
        return;
    }
    const int nworkers = NWARPS;
    const long long wid = (long long)blockIdx.x * NWARPS + warp;
    const long long W = (long long)d.G * NWARPS;
    const uint32_t ring = sh.ring0 + (uint32_t)warp * RING_BYTES_PER_WARP;
    bool prefetched = false;
    long long t_wait = 0, t_rows = 0, t_arrive = 0, n_slow = 0, t_tiles = 0, n_pref = 0, n_notready = 0, n_look = 0, n_look_ok = 0;
    // phase counters (CGG_PROFILE=1) exist only in builds with -DCGG_PROFILE_BUILD: even switched off at run time they held 18
    // registers across the row loops of a kernel that sits at the 255-register limit
#ifdef CGG_PROFILE_BUILD
    const bool prof = d.prof != nullptr;
#else
    constexpr bool prof = false;
#endif
    // ---- wait until decision #round of chain c is published (usually it already is).  One warp at a time polls the
    // flag for the whole CTA and, when it has advanced, fetches the chain's control block with one coalesced request
    // into shared memory; the other warps only watch shared memory.  Returns false if the wait timed out.
    auto wait_decision = [&](int c, unsigned long long round) -> bool {
        volatile unsigned long long *sv = &sh.ver[c];
        if (ver_below(&sh.ver[c], round, lane)) {
            ++n_slow;
            const unsigned long long t0 = globaltimer_ns();
            unsigned spins = 0;
            for (;;) {
                int got = 0;
                if (lane == 0) got = (atomicCAS_block(&sh.lock[c], 0, 1) == 0);
                got = __shfl_sync(0xffffffffu, got, 0);
                if (got) {
                    unsigned long long v = 0;
                    if (lane == 0) v = ld_acquire_u64(&d.sync[c].version);
                    v = __shfl_sync(0xffffffffu, v, 0);
                    if (v < round) ++n_notready;
                    if (__any_sync(0xffffffffu, lane == 0 && v >= round && v > ver_read(&sh.ver[c]))) {
                        if (lane < CTL_WORDS) sh.ctl[c * CTL_WORDS + lane] = __ldcg(reinterpret_cast<const double *>(d.ctl + c) + lane);
                        __syncwarp();
                        if (lane == 0) { __threadfence_block(); *sv = v; }
                    }
                    if (lane == 0) { __threadfence_block(); atomicExch_block(&sh.lock[c], 0); }
                }
                if (!ver_below(&sh.ver[c], round, lane)) break;
                __nanosleep(spins < 8 ? 64 : 256);
                if ((++spins & 63u) == 0 && wait_timed_out(d, t0, lane)) {
                    return false;
                }
            }
        }
        __syncwarp();
        return true;
    };
#ifdef CGG_GROUP_PASSES
    double accg[4][NV];
    double (&acc)[NV] = accg[0];
    double (&acc2)[NV] = accg[1];
#else
    double acc[NV], acc2[NV];
#endif
    bool pair_prefetched = false;
    int pf_j = -1, pf_cj = -1, pf_g = 0; bool pf_full = false;      // what the pair / group prologue issued ahead of time was issued for
    ColCache cc;        // this warp's X-column cache (pair and group passes); lives for the launch
    cc.cap = d.colcache; cc.base = sh.cache0 + (uint32_t)warp * (uint32_t)(2 * d.colcache) * 512u + (uint32_t)lane * 16u;
    cc.tag0 = cc.tag1 = -1; cc.fill_col = cc.fill_slot = -1;
    // group passes (four chains per walk): the steady-state kinds of pass only (binomial light passes, gaussian)
#ifdef CGG_GROUP_PASSES
    constexpr bool GROUP_FAMILY = FAMILY == CGG_BINOMIAL || FAMILY == CGG_GAUSSIAN;
#else
    constexpr bool GROUP_FAMILY = false;      // (group passes are an experiment: build with -DCGG_GROUP_PASSES)
#endif
    int n_group = 0;
    for (unsigned long long round = 0;; ++round) {
        bool any = false;
        for (int c = 0; c < d.C; ++c) {
            long long tA = prof ? clock64() : 0;
            if (prof && lane == 0 && round >= 1 && round <= 128) {
                if (blockIdx.x == 20 && warp == 0) d.prof[32 + 4096 + 32 * 128 * 4 + 2 * 1024 * 32 + (c * 128 + (round - 1))] = globaltimer_ns();
            }
            if (!wait_decision(c, round)) return;
            // ---- chains 2k and 2k + 1 at the same coordinate share one walk over the rows (pair pass)
            bool pair = false;
            if (d.pair && !(c & 1) && c + 1 < d.C &&
                ((unsigned)(__double_as_longlong(sh.ctl[c * CTL_WORDS + 1]) >> 32) & JET_BIT)) {     // only a jet pass can be shared
                if (!wait_decision(c + 1, round)) return;
                // the shared control blocks must be exactly the ones of this round (a finished chain's never are)
                pair = ver_is(&sh.ver[c], round, lane) && ver_is(&sh.ver[c + 1], round, lane) &&
                       pair_batchable(sh.ctl + c * CTL_WORDS, sh.ctl + (c + 1) * CTL_WORDS);
            }
            // ---- ... and chains 4k .. 4k + 3 (group pass) when all four say the same and the X-column cache serves the walk
            int g = pair ? 2 : 1;
            if constexpr (GROUP_FAMILY) if (pair && d.quad && !(c & 3) && c + 3 < d.C && cc.cap > 0) {
                const double *cw = sh.ctl + c * CTL_WORDS;
                const long long w1 = __double_as_longlong(cw[1]);
                const bool full = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                if (FAMILY != CGG_BINOMIAL || !full) {
                    if (!wait_decision(c + 2, round) || !wait_decision(c + 3, round)) return;
                    if (ver_is(&sh.ver[c + 2], round, lane) && ver_is(&sh.ver[c + 3], round, lane) &&
                        pair_batchable(cw, cw + 2 * CTL_WORDS) && pair_batchable(cw, cw + 3 * CTL_WORDS) &&
                        group_cache_ok(cc, (int)(w1 & 0xffffffffLL))) g = 4;
                }
            }
            // kind of walk: 1 single, 2 pair (general: any operand source), 3 lean pair (the steady state: X_commit held by the
            // X-column cache, X_j held or filled by the pass itself -- GroupStream<2>, no operand-source branches), 4 group of four
            int kind = g;
#ifdef CGG_LEAN_PAIR
            constexpr bool LFULL = FAMILY != CGG_BINOMIAL;           // the steady-state kind of jet pass of the family
            if (g == 2 && cc.cap > 0) {
                const long long w1 = __double_as_longlong(sh.ctl[c * CTL_WORDS + 1]);
                const bool full = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                if (full == LFULL && group_cache_ok(cc, (int)(w1 & 0xffffffffLL))) kind = 3;
            }
#endif
            long long tB = prof ? clock64() : 0;
            t_wait += tB - tA;
            if (pair_prefetched) {
                // the first tiles of this pass were requested ahead of time, possibly from a PREDICTED control block (same kind of
                // pass, next column): they are only good if the real block says the same
                bool good = kind >= 2 && kind == pf_g;
                if (good) {
                    const double *cwA = sh.ctl + c * CTL_WORDS;
                    const long long w0 = __double_as_longlong(cwA[0]), w1 = __double_as_longlong(cwA[1]);
                    const bool full = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                    good = (int)(w0 & 0xffffffffLL) == pf_j && (int)(w1 & 0xffffffffLL) == pf_cj && full == pf_full;
                }
                if (!good) { cp_async_wait<0>(); pair_prefetched = false; ++n_notready; }
            }
            if (pair && prefetched) { cp_async_wait<0>(); prefetched = false; }               // a single-pass prefetch of chain c: other ring layout
#ifdef CGG_GROUP_PASSES
            if constexpr (GROUP_FAMILY) if (g == 4) {
                const double *cw0 = sh.ctl + c * CTL_WORDS;
                constexpr bool GFULL = FAMILY != CGG_BINOMIAL;          // (binomial: light passes only, see above)
                const bool was_pref = pair_prefetched;
                pair_prefetched = false;
                // which group comes next, and in which round
                int c2 = -1; unsigned long long nround = round;
                if (c + 4 >= d.C) { c2 = 0; nround = round + 1; }               // this is the last item of the round
                else if (c + 7 < d.C) c2 = c + 4;                               // another group of four follows
                if (c2 == c) c2 = -1;                                           // (the only group: its next decision cannot be there yet)
                auto early = [&]() {      // with three or more groups the next group's decisions are published about a pass ahead
                    if (c2 >= 0 && d.C >= 12 && warp == (int)((round + (unsigned long long)(c >> 2)) % NWARPS)) {
                        pair_lookahead(d, sh, c2, nround, lane); pair_lookahead(d, sh, c2 + 2, nround, lane);
                    }
                };
                auto after = [&]() {
                    if (c2 < 0) return;
                    double pred[2];
                    const double *nA = sh.ctl + c2 * CTL_WORDS;
                    const bool l0 = pair_lookahead(d, sh, c2, nround, lane);
                    const bool l1 = l0 && pair_lookahead(d, sh, c2 + 2, nround, lane);
                    if (l1) {
                        if (!(pair_batchable(nA, nA + CTL_WORDS) && pair_batchable(nA, nA + 2 * CTL_WORDS) && pair_batchable(nA, nA + 3 * CTL_WORDS))) return;
                    } else if (!l0) {
                        // not published yet: predict the pass (same kind, next column, committing the column just sampled), as the
                        // pair passes do; the real blocks are compared with the prediction before the tiles are used
                        // (the block may be rewritten by another warp of the CTA at any moment: one lane reads it for the warp)
                        const long long w0 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[0]), 0), w1 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[1]), 0);
                        const int jp = (int)(w0 & 0xffffffffLL);
                        const unsigned mk = (unsigned)(w1 >> 32);
                        if (!(mk & JET_BIT) || jp < 0 || (long long)jp >= (long long)d.p) return;
                        const int jn = (jp + 1 == (int)d.p) ? 0 : jp + 1;
                        pred[0] = __longlong_as_double((long long)(unsigned)jn);
                        pred[1] = __longlong_as_double(((long long)mk << 32) | (long long)(unsigned)jp);
                        nA = pred;
                    }                                                           // (l0 only: chain c2's real block stands for the group)
                    const long long w0 = __double_as_longlong(nA[0]), w1 = __double_as_longlong(nA[1]);
                    const bool full2 = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                    if (!(((unsigned)(w1 >> 32)) & JET_BIT) || (int)(w0 & 0xffffffffLL) < 0 || full2 != GFULL) return;
                    if (!group_cache_ok(cc, (int)(w1 & 0xffffffffLL))) return;
                    pf_j = (int)(w0 & 0xffffffffLL); pf_cj = (int)(w1 & 0xffffffffLL); pf_full = full2; pf_g = 4;
                    GroupStream<4> ns(d, c2, nA, wid, W, lane, ring, cc, FAMILY != CGG_BINOMIAL || full2);
                    ns.prologue(false);
                    pair_prefetched = true;
                };
                warp_pass_group<FAMILY, GFULL, 4>(d, c, cw0, wid, W, lane, ring, s_l1p, was_pref, cc, early, after, accg);
                any = true;
                ++n_group;
                n_pref += pair_prefetched ? 1 : 0;
                long long tC = prof ? clock64() : 0;
                t_rows += tC - tB; t_tiles += tC - tB;
                const int nvd = jet_nvals(FAMILY, !GFULL);
#pragma unroll
                for (int k = 0; k < 4; ++k) cta_deliver_limbs(d, sh, c + k, nvd, warp, lane, nworkers, accg[k]);
                if (prof) t_arrive += clock64() - tC;
                c += 3;
                continue;
            }
#endif
#ifdef CGG_LEAN_PAIR
            if (kind == 3) {
                const double *cw0 = sh.ctl + c * CTL_WORDS;
                const bool was_pref = pair_prefetched;
                pair_prefetched = false;
                int c2 = -1; unsigned long long nround = round;
                if (d.C >= 6) {                   // (as for general pair passes: with one or two pairs the look would only cost)
                    if (c + 2 >= d.C) { c2 = 0; nround = round + 1; }
                    else if (c + 3 < d.C) { c2 = c + 2; }
                    if (c2 == c) c2 = -1;
                }
                auto early = [&]() {
                    if (c2 >= 0 && warp == (int)((round + (unsigned long long)(c >> 1)) % NWARPS)) pair_lookahead(d, sh, c2, nround, lane);
                };
                auto after = [&]() {
                    if (c2 < 0) return;
                    double pred[2];
                    const double *nA = sh.ctl + c2 * CTL_WORDS;
                    if (pair_lookahead(d, sh, c2, nround, lane)) {
                        if (!pair_batchable(nA, nA + CTL_WORDS)) return;
                    } else {
                        const long long w0 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[0]), 0), w1 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[1]), 0);
                        const int jp = (int)(w0 & 0xffffffffLL);
                        const unsigned mk = (unsigned)(w1 >> 32);
                        if (!(mk & JET_BIT) || jp < 0 || (long long)jp >= (long long)d.p) return;
                        const int jn = (jp + 1 == (int)d.p) ? 0 : jp + 1;
                        pred[0] = __longlong_as_double((long long)(unsigned)jn);
                        pred[1] = __longlong_as_double(((long long)mk << 32) | (long long)(unsigned)jp);
                        nA = pred;
                    }
                    const long long w0 = __double_as_longlong(nA[0]), w1 = __double_as_longlong(nA[1]);
                    const bool full2 = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                    if (!(((unsigned)(w1 >> 32)) & JET_BIT) || (int)(w0 & 0xffffffffLL) < 0 || full2 != LFULL) return;
                    if (!group_cache_ok(cc, (int)(w1 & 0xffffffffLL))) return;
                    pf_j = (int)(w0 & 0xffffffffLL); pf_cj = (int)(w1 & 0xffffffffLL); pf_full = full2; pf_g = 3;
                    GroupStream<2> ns(d, c2, nA, wid, W, lane, ring, cc, FAMILY != CGG_BINOMIAL || full2);
                    ns.prologue(false);
                    pair_prefetched = true;
                };
                double accl[2][NV];
                warp_pass_group<FAMILY, LFULL, 2>(d, c, cw0, wid, W, lane, ring, s_l1p, was_pref, cc, early, after, accl);
                any = true;
                n_pref += pair_prefetched ? 1 : 0;
                long long tC = prof ? clock64() : 0;
                t_rows += tC - tB; t_tiles += tC - tB;
                const int nvd = jet_nvals(FAMILY, !LFULL);
                cta_deliver_limbs(d, sh, c, nvd, warp, lane, nworkers, accl[0]);
                cta_deliver_limbs(d, sh, c + 1, nvd, warp, lane, nworkers, accl[1]);
                if (prof) t_arrive += clock64() - tC;
                ++c;
                continue;
            }
#endif
            if (pair) {
                const double *cwA = sh.ctl + c * CTL_WORDS, *cwB = cwA + CTL_WORDS;
                const bool full = FAMILY != CGG_BINOMIAL || (((unsigned)(__double_as_longlong(cwA[1]) >> 32)) & JET_FULL);
                // when the tiles are consumed: if the NEXT pair's decisions are published and it is a pair pass again,
                // request its first tiles now, before this pair's sums are reduced and delivered
                const bool was_pref = pair_prefetched;
                pair_prefetched = false;
                // which pair comes next, and in which round
                int c2 = -1; unsigned long long nround = round;
                if (d.C >= 6) {                   // with one or two pairs the next decision is never there yet: the look would only cost
                    if (c + 2 >= d.C) { c2 = 0; nround = round + 1; }           // this is the last item of the round
                    else if (c + 3 < d.C) { c2 = c + 2; }                      // another pair follows
                    if (c2 == c) c2 = -1;                                      // (a single chain follows: it needs the ring)
                }
                // One warp of the CTA (a different one every pass) fetches the next pair's control blocks right after its own
                // first tiles were requested -- its round trips to L2 overlap the tiles' -- so that by the time the CTA's
                // warps finish this pair, the blocks are in shared memory and every warp can request the next pair's first
                // tiles at once (a look at the END of the pass only served the warps that finished after the fetch: 25 %).
                auto early = [&]() {
                    if (c2 >= 0 && warp == (int)((round + (unsigned long long)(c >> 1)) % NWARPS)) pair_lookahead(d, sh, c2, nround, lane);
                };
                auto after = [&]() {
                    if (c2 < 0) return;
                    double pred[2];
                    const double *nA = sh.ctl + c2 * CTL_WORDS;
                    if (pair_lookahead(d, sh, c2, nround, lane)) {
                        if (!pair_batchable(nA, nA + CTL_WORDS)) return;
                    } else {
                        // The next pair's decision is not published yet (this CTA is ahead of the slowest ones).  Its first
                        // tiles do not depend on the decision's VALUE, only on which pass comes: in the stationary regime
                        // that is a jet pass of the same kind on the next column, applying the update of the column just
                        // sampled.  Request them from that prediction; the real block is compared with it before use.
                        // (the block may be rewritten by another warp of the CTA at any moment: one lane reads it for the warp)
                        const long long w0 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[0]), 0), w1 = __shfl_sync(0xffffffffu, __double_as_longlong(nA[1]), 0);
                        const int jp = (int)(w0 & 0xffffffffLL);
                        const unsigned mk = (unsigned)(w1 >> 32);
                        if (!(mk & JET_BIT) || jp < 0 || (long long)jp >= (long long)d.p) return;
                        const int jn = (jp + 1 == (int)d.p) ? 0 : jp + 1;
                        pred[0] = __longlong_as_double((long long)(unsigned)jn);                                  // ncand = 0
                        pred[1] = __longlong_as_double(((long long)mk << 32) | (long long)(unsigned)jp);            // commit column = the one just sampled
                        nA = pred;
                    }
                    const long long w0 = __double_as_longlong(nA[0]), w1 = __double_as_longlong(nA[1]);
                    const bool full2 = FAMILY != CGG_BINOMIAL || (((unsigned)(w1 >> 32)) & JET_FULL);
                    pf_j = (int)(w0 & 0xffffffffLL); pf_cj = (int)(w1 & 0xffffffffLL); pf_full = full2; pf_g = 2;
                    PairStream ns(d, c2, nA, wid, W, lane, ring, cc, full2);
                    ns.prologue(false);
                    pair_prefetched = true;
                };
                if (full) warp_pass_jet2<FAMILY, true>(d, c, cwA, cwB, wid, W, lane, ring, s_l1p, was_pref, cc, early, after, acc, acc2);
                else warp_pass_jet2<FAMILY, false>(d, c, cwA, cwB, wid, W, lane, ring, s_l1p, was_pref, cc, early, after, acc, acc2);
                any = true;
                n_pref += pair_prefetched ? 1 : 0;
                long long tC = prof ? clock64() : 0;
                t_rows += tC - tB; t_tiles += tC - tB;
                const int nvd = jet_nvals(FAMILY, !full);
                cta_deliver_limbs(d, sh, c, nvd, warp, lane, nworkers, acc);
                cta_deliver_limbs(d, sh, c + 1, nvd, warp, lane, nworkers, acc2);
                if (prof) t_arrive += clock64() - tC;
                ++c;
                continue;
            }
            // ---- stream this warp's rows for chain c; if the next chain's decision is already published, its
            // first tiles are requested as soon as this chain's tiles are consumed (cross-chain prefetch)
            int j = -1;
            const int nxt = (c + 1 == d.C) ? 0 : c + 1;
            const unsigned long long nround = (c + 1 == d.C) ? round + 1 : round;
            LookAhead la{&sh, d.sync, d.ctl, d.C > 1 ? nxt : -1, nround, 0, 0};
            const int nc = worker_pass<FAMILY>(d, c, sh.ctl + c * CTL_WORDS, wid, W, lane, ring, s_l1p, sh.sacc + warp * KMAX * 32,
                                               acc, j, prefetched, nxt, &la, prof ? &t_tiles : nullptr);
            n_look += la.n_look; n_look_ok += la.n_ok;
            if (nc < 0) continue;
            any = true;
            long long tC = prof ? clock64() : 0;
            t_rows += tC - tB;
            // ---- CTA-level then grid-level arrival
            cta_deliver_limbs(d, sh, c, nc, warp, lane, nworkers, acc);
            if (prof && lane == 0 && round == 60 && c == 0)
                d.prof[32 + 4096 + 32 * 128 * 4 + (blockIdx.x * NWARPS + warp) * 2 + 1] = globaltimer_ns();
            if (prof && lane == 0 && round < 128) {     // arrival spread of the warps for pass #round of chain c
                const unsigned long long tn = globaltimer_ns();
                atomicMin(d.prof + 32 + 4096 + (c * 128 + round) * 4 + 2, tn);
                atomicMax(d.prof + 32 + 4096 + (c * 128 + round) * 4 + 3, tn);
            }
            if (prof) t_arrive += clock64() - tC;
        }
        if (!any) break;
    }
    if (n_group && wid == 0 && lane == 0) atomicAdd(&d.hdr->group_passes, n_group);
    if (prof && lane == 0) {
        atomicAdd(d.prof + 0, (unsigned long long)t_wait); atomicAdd(d.prof + 1, (unsigned long long)t_rows);
        atomicAdd(d.prof + 2, (unsigned long long)t_arrive); atomicAdd(d.prof + 5, (unsigned long long)n_slow);
        atomicAdd(d.prof + 3, (unsigned long long)t_tiles); atomicAdd(d.prof + 4, (unsigned long long)n_pref);
        atomicAdd(d.prof + 6, 1ULL);
        atomicAdd(d.prof + 9, (unsigned long long)n_notready); atomicAdd(d.prof + 10, (unsigned long long)n_look); atomicAdd(d.prof + 11, (unsigned long long)n_look_ok);
        // per-CTA view (who waits, who never does): [16 + 4 * cta + {wait, tiles, rows, smid}]
        unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicAdd(d.prof + 32 + 4 * blockIdx.x + 0, (unsigned long long)t_wait);
        atomicAdd(d.prof + 32 + 4 * blockIdx.x + 1, (unsigned long long)t_tiles);
        atomicAdd(d.prof + 32 + 4 * blockIdx.x + 2, (unsigned long long)t_rows);
        d.prof[32 + 4 * blockIdx.x + 3] = smid;
        // per-warp view: [16 + 4096 + 32*128*4 + (cta * NWARPS + warp) * 2 + {wait, tiles}]
        d.prof[32 + 4096 + 32 * 128 * 4 + (blockIdx.x * NWARPS + warp) * 2 + 0] = (unsigned long long)t_wait;
    }
}

// K3s (cluster driver, small n).  One thread-block CLUSTER per chain, no communication through global memory at all: the
// S CTAs of a cluster split the chain's rows (the same tile -> warp ownership as everywhere else), every CTA sends its
// sums of a pass into the shared memory of ALL CTAs of its cluster (distributed shared memory), one cluster barrier, and
// then every CTA adds the S parts in rank order and runs the chain's decision ITSELF -- S identical, deterministic copies
// of the state machine instead of one decision that has to be published, polled for and fetched.  With n = 1e5 a pass is
// ~1-2 us of streaming, so the grid-wide protocol of the persistent driver (decider CTA, limb accumulators, version flags:
// ~10 us per update and chain) is all latency; here an update is pass + barrier + decision.  The global side effects of a
// decision (beta, slice-width estimate, sample store) are written by every CTA with identical values, so each CTA reads
// back what it wrote itself.  Chains do not interact: C clusters run concurrently (or in waves), no cooperative launch.
constexpr int CLUSTER_MAX = 16;
struct ClusterShared {
    double *sacc, *part, *xs, *ctl, *vals;
    DeciderCache *dc;
    uint32_t ring0;
    __device__ __forceinline__ ClusterShared(unsigned char *base) {
        ring0 = (uint32_t)__cvta_generic_to_shared(base);
        sacc = reinterpret_cast<double *>(base + NWARPS * RING_BYTES_PER_WARP);
        part = sacc + NWARPS * KMAX * 32;              // [NWARPS][NV] the warps' partial sums
        xs = part + NWARPS * NV;                       // [2][CLUSTER_MAX][NV] every CTA's sums of the pass, double-buffered by pass parity
        ctl = xs + 2 * CLUSTER_MAX * NV;               // [CTL_WORDS] the control block of the coming pass
        vals = ctl + CTL_WORDS;                        // [NV] the pass's totals
        dc = reinterpret_cast<DeciderCache *>((reinterpret_cast<uintptr_t>(vals + NV) + 127) & ~(uintptr_t)127);
    }
    static size_t bytes() {
        return (size_t)NWARPS * RING_BYTES_PER_WARP + sizeof(double) * (NWARPS * KMAX * 32 + NWARPS * NV + 2 * CLUSTER_MAX * NV + CTL_WORDS + NV + 2) + sizeof(DeciderCache) + 256;
    }
};
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(uint32_t local_smem_addr, unsigned rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

template <int FAMILY>
__global__ void __launch_bounds__(THREADS, 1) sweep_cluster_kernel(const __grid_constant__ Dev d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double2 s_l1p[MATH_TAB_N];
    ClusterShared sh(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int c = (int)blockIdx.x / S;                 // one cluster per chain
    load_l1p_table(s_l1p);
    for (int i = threadIdx.x; i < CTL_WORDS; i += THREADS) sh.ctl[i] = __ldcg(reinterpret_cast<const double *>(d.ctl + c) + i);
    if (threadIdx.x == 0) { sh.dc->valid = 0; sh.dc->pref_j = -1; sh.dc->pre.valid = 0; }
    __syncthreads();
    // warp 0 is the CTA's decider (it prepares the coming decision while the other warps stream the rows); warps 1.. are the workers
    constexpr int NWORK = NWARPS - 1;
    const long long wid = (long long)rank * NWORK + (warp - 1), W = (long long)S * NWORK;
    const uint32_t ring = sh.ring0 + (uint32_t)warp * RING_BYTES_PER_WARP;
    const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(sh.xs);
    double acc[NV];
#ifdef CGG_PROFILE_BUILD
    const bool prof = d.prof != nullptr && blockIdx.x == 0 && lane == 0;      // CGG_PROFILE: phase timers of the first CTA
#else
    constexpr bool prof = false;
#endif
    for (unsigned long long pass = 0;; ++pass) {
        int nc = 0;
        const long long tp0 = prof ? clock64() : 0;
        if (warp > 0) {
            int j;
            bool prefetched = false;
            NoLookAhead nola;
            nc = worker_pass<FAMILY>(d, c, sh.ctl, wid, W, lane, ring, s_l1p, sh.sacc + warp * KMAX * 32, acc, j, prefetched, 0, &nola);
            if (prof && warp == 1) atomicAdd(d.prof + 22, (unsigned long long)(clock64() - tp0));
            if (nc > 0 && lane == 0) {
#pragma unroll
                for (int k = 0; k < NV; ++k) if (k < nc) sh.part[warp * NV + k] = acc[k];
            }
        } else {
            // what worker_pass will return for this control block (the decider does not stream)
            const long long w0 = __double_as_longlong(sh.ctl[0]), w1 = __double_as_longlong(sh.ctl[1]);
            const int j = (int)(w0 & 0xffffffffLL), ncand = (int)(w0 >> 32), cj = (int)(w1 & 0xffffffffLL);
            const unsigned cmask = (unsigned)(w1 >> 32);
            nc = (j < 0) ? -1 : ((cmask & JET_BIT) ? jet_nvals(FAMILY, !(cmask & JET_FULL)) : ((ncand == 0 && cj < 0) ? 0 : ncand));
            if (nc >= 0 && sh.dc->valid) { decider_prefetch(d, c, sh.dc, lane); decider_prephase(d, c, sh.dc, lane); }
            if (prof) { atomicAdd(d.prof + 20, (unsigned long long)(clock64() - tp0)); atomicAdd(d.prof + 24, 1ULL); }
        }
        if (nc < 0) break;                             // the chain has finished (every thread of the cluster sees the same block)
        __syncthreads();
        // CTA fold in warp order, then the CTA's sums go to slot [parity][rank] of every CTA of the cluster
        const int par = (int)(pass & 1ULL);
        if (warp == 0 && lane < nc) {
            double v = 0.0;
#pragma unroll
            for (int w = 1; w < NWARPS; ++w) v += sh.part[w * NV + lane];
            for (int r = 0; r < S; ++r) st_cluster_f64(xs_addr + (uint32_t)(((par * CLUSTER_MAX + rank) * NV + lane) * 8), (unsigned)r, v);
        }
        const long long tb0 = prof ? clock64() : 0;
        cluster_barrier();
        if (warp == 0) {
            const long long td0 = prof ? clock64() : 0;
            if (lane < NV) {
                double v = 0.0;
                if (lane < nc) for (int r = 0; r < S; ++r) v += sh.xs[(par * CLUSTER_MAX + r) * NV + lane];       // rank order: identical in every CTA
                sh.vals[lane] = v;
            }
            __syncwarp();
            decide_chain(&d, c, lane, -1, SRC_SLOTS, sh.dc, sh.vals);
            if (lane < CTL_WORDS) sh.ctl[lane] = reinterpret_cast<const double *>(&sh.dc->ct)[lane];
            __syncwarp();
            if (prof) { atomicAdd(d.prof + 21, (unsigned long long)(clock64() - td0)); atomicAdd(d.prof + 23, (unsigned long long)(td0 - tb0)); }
        }
        __syncthreads();
    }
    cluster_barrier();      // nobody leaves while a peer may still write into its shared memory
}

// Stepwise driver: one pass of every chain per launch; the last CTA to finish decides.
// mode 0: run the state machines.  mode 1 (row-sharded / operator): only fold the exact accumulators
// into d.xbuf (+ ll_const) for the exchange / the caller.
template <int FAMILY>
__global__ void __launch_bounds__(THREADS, 1) pass_kernel(const __grid_constant__ Dev d, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared sh(smem_raw, d.C);
    __shared__ int s_flag;
    __shared__ double2 s_l1p[MATH_TAB_N];
    if (d.hdr->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_l1p_table(s_l1p);
    for (int i = threadIdx.x; i < d.C; i += THREADS) sh.cnt[i] = 0;
    for (int i = threadIdx.x; i < d.C * CTL_WORDS; i += THREADS) sh.ctl[i] = __ldcg(reinterpret_cast<const double *>(d.ctl) + i);
    __syncthreads();
    const long long wid = (long long)blockIdx.x * NWARPS + warp, W = (long long)d.G * NWARPS;
    const uint32_t ring = sh.ring0 + (uint32_t)warp * RING_BYTES_PER_WARP;
    double acc[NV];
    for (int c = 0; c < d.C; ++c) {
        int j;
        bool prefetched = false;
        NoLookAhead nola;
        const int nc = worker_pass<FAMILY>(d, c, sh.ctl + c * CTL_WORDS, wid, W, lane, ring, s_l1p, sh.sacc + warp * KMAX * 32,
                                           acc, j, prefetched, 0, &nola);
        if (nc > 0) cta_deliver(d, sh, c, nc, warp, lane, NWARPS, acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&d.hdr->ticket, 1u);
        s_flag = (t == (unsigned)d.G - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    for (int c = warp; c < d.C; c += NWARPS) {
        if (mode == 0) decide_chain(&d, c, lane, -1, SRC_ACC);
        else {
            const unsigned cm = (unsigned)d.ctl[c].coarse_mask;
            const bool jet = (cm & JET_BIT) != 0u, full = (cm & JET_FULL) != 0u;
            const int nc = jet ? jet_nvals(d.family, !full) : d.ctl[c].ncand;
            // the additive constant goes with every candidate sum and, row-sharded, with the M_0 of a full jet pass (value 0)
            const bool with_const = !jet || (d.sharded && lane == 0 && (full || d.family != CGG_BINOMIAL));
            if (lane < NV) d.xbuf[c * NV + lane] = (lane < nc) ? acc_take(d.acc + c * NV + lane) + (with_const ? d.ll_const : 0.0) : 0.0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int done = 1;
        for (int c = 0; c < d.C; ++c) done &= (d.ctl[c].j < 0);
        if (mode == 0 && done) d.hdr->done = 1;
        d.hdr->ticket = 0;
    }
}

// Row-sharded mode: the state machines alone, after the exchange made d.xbuf global sums.
__global__ void __launch_bounds__(THREADS, 1) decide_kernel(const __grid_constant__ Dev d) {
    if (d.hdr->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < d.C; c += NWARPS) decide_chain(&d, c, lane, -1, SRC_XBUF);
    __syncthreads();
    if (threadIdx.x == 0) {
        int done = 1;
        for (int c = 0; c < d.C; ++c) done &= (d.ctl[c].j < 0);
        if (done) d.hdr->done = 1;
    }
}

// cgg_log_potential tail: out[k] = ll_k + sum_{l != j} log pi(beta_l) + log pi(cand_k)
__global__ void finalize_eval_kernel(Dev d, int c, int j, int K, const double *cand, double *out, double *prior_sum_out) {
    const int lane = threadIdx.x;
    double v = 0.0, vall = 0.0;
    for (int64_t l = lane; l < d.p; l += 32) {
        const double t = prior_logdens(d.prior, d.beta[(int64_t)c * d.p + l]);
        vall += t;
        if (l != j) v += t;
    }
    v = warp_sum(v);
    vall = warp_sum(vall);
    if (lane < K) out[lane] = d.xbuf[c * NV + lane] + (v + prior_logdens(d.prior, cand[lane]));
    if (lane == 0 && prior_sum_out) *prior_sum_out = vall;
}

// K4: eta = X %*% beta (R/mcmcglm.R:215).  Column-ordered accumulation with separate multiply and
// add so the result is bit-identical to oracle.c:orc_init_eta.
__global__ void __launch_bounds__(THREADS) init_eta_kernel(Dev d, int c) {
    const double *beta = d.beta + (int64_t)c * d.p;
    double *eta = d.eta + (int64_t)c * d.lde;
    for (int64_t i = 2 * ((int64_t)blockIdx.x * THREADS + threadIdx.x); i < d.n; i += 2LL * gridDim.x * THREADS) {
        if (i + 1 < d.n) {
            double2 a = make_double2(0.0, 0.0);
            for (int64_t l = 0; l < d.p; ++l) {
                const double2 x = ld_stream2(d.X + l * d.ldx + i);
                const double bl = beta[l];
                a.x = __dadd_rn(a.x, __dmul_rn(x.x, bl));
                a.y = __dadd_rn(a.y, __dmul_rn(x.y, bl));
            }
            *reinterpret_cast<double2 *>(eta + i) = a;
        } else {
            double a = 0.0;
            for (int64_t l = 0; l < d.p; ++l) a = __dadd_rn(a, __dmul_rn(d.X[l * d.ldx + i], beta[l]));
            eta[i] = a;
        }
    }
}

// linear_predictor_calc = "naive" (R/glm_utils.R:206-208): the reference then forms new_eta <- X %*% new_beta for every
// evaluation, O(n p) instead of the O(n) update.  The engine's naive mode recomputes eta = X beta from scratch (K4) before
// every pass over the rows; an accepted value is already in beta, so the deferred eta update of that chain is dropped.
__global__ void naive_drop_commit_kernel(Dev d) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    d.ctl[c].commit_j = -1; d.ctl[c].commit_delta = 0.0;
}

// K2 stand-alone: eta <- eta + X_j * diff (R/glm_utils.R:126-132)
__global__ void __launch_bounds__(THREADS) axpy_eta_kernel(Dev d, int c, int64_t j, double diff) {
    const double *xj = d.X + j * d.ldx;
    double *eta = d.eta + (int64_t)c * d.lde;
    for (int64_t i = 2 * ((int64_t)blockIdx.x * THREADS + threadIdx.x); i < d.n; i += 2LL * gridDim.x * THREADS) {
        if (i + 1 < d.n) {
            double2 e = *reinterpret_cast<double2 *>(eta + i);
            const double2 x = ld_stream2(xj + i);
            e.x = eta_shift(e.x, x.x, diff);
            e.y = eta_shift(e.y, x.y, diff);
            *reinterpret_cast<double2 *>(eta + i) = e;
        } else {
            eta[i] = eta_shift(eta[i], xj[i], diff);
        }
    }
}

// Diagnostic: the per-row log-density term (without the per-dataset constant) of each (y_i, eta_i).
__global__ void row_terms_kernel(int family, int64_t n, const double *y, const double *eta, double inv_sd, double *out) {
    __shared__ double2 s_l1p[MATH_TAB_N];
    load_l1p_table(s_l1p);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // scored through the same two-row code path as the sweep kernels (the row duplicated, x = 0)
        const double2 yy = make_double2(y[i], y[i]), ee = make_double2(eta[i], eta[i]), xx = make_double2(0.0, 0.0);
        double v;
        if (family == CGG_GAUSSIAN) v = RowPair<CGG_GAUSSIAN>(yy, ee, xx).term(0.0, inv_sd, s_l1p);
        else if (family == CGG_BINOMIAL) v = RowPair<CGG_BINOMIAL>(yy, ee, xx).term(0.0, inv_sd, s_l1p);
        else if (family == CGG_KF_NEGBIN) v = RowPair<CGG_KF_NEGBIN>(yy, ee, xx).term(0.0, inv_sd, s_l1p);
        else if (family == CGG_KF_PROBIT) v = RowPair<CGG_KF_PROBIT>(yy, ee, xx).term(0.0, inv_sd, s_l1p);
        else v = RowPair<CGG_POISSON>(yy, ee, xx).term(0.0, inv_sd, s_l1p);
        out[i] = 0.5 * v;
    }
}

// Diagnostic: max over ALL fp32 s with |s| <= 37 of |softplus32(s) - softplus(s)| / (1 + |s|), the constant the
// coarse pre-filter's error bound is built on.  out[0] = that maximum, out[1] = the s where it occurs.
__global__ void coarse_error_scan_kernel(double *out, unsigned long long *out_bits) {
    __shared__ double2 s_l1p[MATH_TAB_N];
    load_l1p_table(s_l1p);
    __syncthreads();
    double worst = 0.0;
    unsigned int worst_bits = 0;
    const unsigned int top = __float_as_uint(37.0f);
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= top; b += (unsigned long long)gridDim.x * blockDim.x) {
        for (int sign = 0; sign < 2; ++sign) {
            const float s = __uint_as_float((unsigned int)b | (sign ? 0x80000000u : 0u));
            if (fabsf(fabsf(s) - 30.0f) < 0.02f) continue;           // excluded from the bound by construction
            bool near = false;
            const float got = softplus32(s, near);
            double o0, o1;
            softplus2((double)s, (double)s, s_l1p, o0, o1);
            const double e = fabs((double)got - o0) / (1.0 + fabs((double)s));
            if (e > worst) { worst = e; worst_bits = __float_as_uint(s); }
        }
    }
    // block max
    __shared__ double s_w[32]; __shared__ unsigned int s_b[32];
    for (int o = 16; o > 0; o >>= 1) {
        const double w2 = __shfl_xor_sync(0xffffffffu, worst, o); const unsigned int b2 = __shfl_xor_sync(0xffffffffu, worst_bits, o);
        if (w2 > worst) { worst = w2; worst_bits = b2; }
    }
    if ((threadIdx.x & 31) == 0) { s_w[threadIdx.x >> 5] = worst; s_b[threadIdx.x >> 5] = worst_bits; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) if (s_w[i] > worst) { worst = s_w[i]; worst_bits = s_b[i]; }
        out[blockIdx.x] = worst; out_bits[blockIdx.x] = worst_bits;
    }
}

// Diagnostic: max absolute error of the binomial LIGHT jet row's quantities (ua = tanh(a/2), v = s(1-s), v ua) against
// libdevice over a grid of a = |eta| in [0, 40): the constant JET_LIGHT_EPS (cgg_jet.cuh) has to cover it.
__global__ void light_error_scan_kernel(double *out, int per_thread) {
    __shared__ double2 s_l1p[MATH_TAB_N];
    __shared__ double s_w[32];
    load_l1p_table(s_l1p);
    __syncthreads();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    double worst = 0.0;
    for (int i = 0; i < per_thread; ++i) {
        const double a = 40.0 * ((double)(tid + (long long)i * nthr) + 0.37) / (double)(nthr * per_thread);
        for (int sgn = 0; sgn < 2; ++sgn) {
            double m[NV] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            unsigned risk = 0;
            JetRow<CGG_BINOMIAL>::add_light(sgn ? -a : a, 1.0, s_l1p, m, risk);
            const double sg = 1.0 / (1.0 + exp(-a)), ua = 2.0 * sg - 1.0, v = sg * (1.0 - sg);
            const double xh = sgn ? -1.0 : 1.0;        // xh = sign(eta) xs: m1 = xh ua, m2 = v, m3 = xh v ua
            worst = fmax(worst, fmax(fabs(m[1] - xh * ua), fmax(fabs(m[2] - v), fabs(m[3] - xh * v * ua))));
        }
    }
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = worst;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) worst = fmax(worst, s_w[i]);
        out[blockIdx.x] = worst;
    }
}

// One-time scan of y: support check per family, and sum lgamma(y + 1) for poisson.
__global__ void __launch_bounds__(THREADS) scan_y_kernel(Dev d, double *partial, int *bad) {
    __shared__ double s_red[NWARPS];
    double acc = 0.0;
    int nbad = 0;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * THREADS) {
        const double y = d.y[i];
        if (d.family == CGG_GAUSSIAN) nbad += !isfinite(y);
        else if (d.family == CGG_BINOMIAL || d.family == CGG_KF_PROBIT) nbad += !(y == 0.0 || y == 1.0);
        else {       // poisson, negative binomial: counts
            const bool ok = isfinite(y) && y >= 0.0 && y == floor(y);
            nbad += !ok;
            if (ok && d.family == CGG_POISSON) acc += lgamma(y + 1.0);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    if (nbad) atomicAdd(bad, nbad);
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < NWARPS) ? s_red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

// Per-column statistics for the jet passes (cgg_jet.cuh), one CTA per column, in two steps so that a row-sharded
// handle can reduce them over the ranks in between:
//   col_max_kernel   the binary exponent e of max|x| (frexp), as a one-hot entry of a small histogram per column
//                    (bins for e in [-CS_EMAX, CS_EMAX]; anything outside, or a non-finite entry, goes to the two edge
//                    bins = "this column cannot be scaled").  Histograms ADD across ranks; the top non-empty bin is the
//                    global exponent, so a sum-only exchange is enough.
//   col_sums_kernel  cs = 2^-e (max|x| * cs in [0.5, 1): exact scaling), S_k = sum_i |x_i cs|^k for k = 1..8 and, for the
//                    binomial family, C1 = sum_i x_i cs (y_i - 1/2) with compensated (two-sum) accumulation.  S_k and C1
//                    add across ranks; col_finish_kernel rounds S_k UP (they enter error bounds).
constexpr int CS_EMAX = 64;
constexpr int CS_BINS = 2 * CS_EMAX + 1;
__device__ __forceinline__ void two_sum_acc(double &s, double &c, double v) {
    const double t = s + v;
    const double bp = t - s;
    c += (s - (t - bp)) + (v - bp);
    s = t;
}
__global__ void __launch_bounds__(THREADS) col_max_kernel(const double *X, int64_t n, int64_t ldx, int64_t p, double *hist /* [p][CS_BINS], zeroed */) {
    __shared__ double s_red[NWARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t j = blockIdx.x; j < p; j += gridDim.x) {
        const double *x = X + j * ldx;
        double mx = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += THREADS) { const double a = fabs(x[i]); mx = (a > mx || a != a) ? a : mx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, mx, o); mx = (t > mx || t != t) ? t : mx; }
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int w = 0; w < NWARPS; ++w) { const double t = s_red[w]; m = (t > m || t != t) ? t : m; }
            int bin = CS_EMAX;                                // an all-zero column: exponent 0, cs = 1
            if (m > 0.0 && m < INFINITY) {
                int e = 0;
                frexp(m, &e);
                bin = (e < -CS_EMAX + 1 || e > CS_EMAX - 1) ? (e < 0 ? 0 : CS_BINS - 1) : e + CS_EMAX;
            } else if (m != 0.0) bin = CS_BINS - 1;           // Inf / NaN
            hist[j * CS_BINS + bin] += 1.0;
        }
        __syncthreads();
    }
}
// cs of column j from its (globally summed) exponent histogram; 0 = the column cannot be scaled
__device__ __forceinline__ double col_scale(const double *hist, int64_t j) {
    int top = -1;
    for (int b = CS_BINS - 1; b >= 0; --b) if (hist[j * CS_BINS + b] != 0.0) { top = b; break; }
    if (top <= 0 || top == CS_BINS - 1) return 0.0;
    return ldexp(1.0, -(top - CS_EMAX));
}
__global__ void __launch_bounds__(THREADS) col_sums_kernel(const double *X, const double *y, int family, int64_t n, int64_t ldx, int64_t p,
                                                           const double *hist, double *sums /* [p][10]: S_1..S_8, C1 sum, C1 compensation */) {
    __shared__ double s_red[NWARPS][10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t j = blockIdx.x; j < p; j += gridDim.x) {
        const double *x = X + j * ldx;
        double cs = col_scale(hist, j);
        if (cs == 0.0) cs = 1.0;
        double S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double c1s = 0.0, c1c = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += THREADS) {
            const double xs = x[i] * cs;
            const double a = fabs(xs);
            double pw = a;
#pragma unroll
            for (int k = 0; k < 8; ++k) { S[k] += pw; pw *= a; }
            if (family == CGG_BINOMIAL) two_sum_acc(c1s, c1c, xs * (y[i] - 0.5));     // y - 1/2 = +-1/2: the product is exact
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { S[k] = warp_sum(S[k]); if (lane == 0) s_red[warp][k] = S[k]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double s2 = __shfl_xor_sync(0xffffffffu, c1s, o), c2 = __shfl_xor_sync(0xffffffffu, c1c, o);
            c1c += c2;
            two_sum_acc(c1s, c1c, s2);
        }
        if (lane == 0) { s_red[warp][8] = c1s; s_red[warp][9] = c1c; }
        __syncthreads();
        if (threadIdx.x < 8) {
            double v = 0.0;
            for (int w = 0; w < NWARPS; ++w) v += s_red[w][threadIdx.x];
            sums[j * 10 + threadIdx.x] = v;
        }
        if (threadIdx.x == 8) {
            double ss = 0.0, cc = 0.0;
            for (int w = 0; w < NWARPS; ++w) { cc += s_red[w][9]; two_sum_acc(ss, cc, s_red[w][8]); }
            sums[j * 10 + 8] = ss; sums[j * 10 + 9] = cc;
        }
        __syncthreads();
    }
}
__global__ void col_finish_kernel(const double *hist, const double *sums, int64_t p, double n_total, double *out /* [p][CS_STRIDE] */) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double cs = col_scale(hist, j);
    double *o = out + j * CS_STRIDE;
    if (cs == 0.0) {            // not scalable: bounds that never decide
        o[0] = 1.0; o[1] = 1.0;
        for (int k = 0; k < 8; ++k) o[2 + k] = INFINITY;
        o[10] = INFINITY; o[11] = 0.0;
        return;
    }
    o[0] = cs; o[1] = 1.0 / cs;
    // fp64 summation of n non-negative terms (and of the per-rank partial sums): relative error <= ~n * 2^-53; inflate
    for (int k = 0; k < 8; ++k) o[2 + k] = sums[j * 10 + k] * (1.0 + 4.0 * n_total * 1.1102230246251565e-16 + 1e-12);
    o[10] = 1.0 / cs;           // >= max|x|
    o[11] = sums[j * 10 + 8] + sums[j * 10 + 9];
}

// Diagnostic (cgg_debug_jet): evaluate the enclosure of chain c's current jet sums at K candidates.
__global__ void jet_debug_kernel(Dev d, int c, int j, int K, int light, double fmag_light, const double *cand,
                                 double *out /* [K] value, [K] bound, [NV] sums */) {
    const int lane = threadIdx.x;
    double m[NV];
    const double mv = (lane < NV) ? d.xbuf[c * NV + lane] : 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) m[k] = __shfl_sync(0xffffffffu, mv, k);
    if (light && d.family == CGG_BINOMIAL) jet_light_unpack(m);
    const double x0 = d.beta[(int64_t)c * d.p + j];
    if (lane < K) {
        double B;
        const double fmag = light ? fmag_light : fabs(m[0]);
        const double dl = jet_eval(d.family, m, d.colstat + (int64_t)j * CS_STRIDE, d.n_total, d.inv_sd, __dadd_rn(cand[lane], -x0), fmag, B, light != 0, d.jet_ce);
        out[lane] = light ? dl : (m[0] + dl) + (d.sharded ? 0.0 : d.ll_const);
        out[K + lane] = B * d.jet_bscale + 8.0 * JET_EPS * (fmag + fabs(dl));
    }
    if (lane < NV) out[2 * K + lane] = m[lane];
}

// Persistent driver, between two launches of one cgg_run: every chain stopped at an iteration boundary (its pending eta
// update flushed); put the running ones back to "start the next coordinate", idle control blocks, versions from zero.
__global__ void rearm_kernel(Dev d) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    Ctl ct;
    memset(&ct, 0, sizeof ct);
    ct.commit_j = -1;
    const bool alive = d.cs[c].status == CGG_OK && !d.hdr->abort;
    if (alive) d.cs[c].phase = PH_START; else ct.j = -1;
    d.ctl[c] = ct;
    d.sync[c].arrive = 0ULL;
    d.sync[c].version = alive ? 0ULL : (1ULL << 62);
}

// Row-sharded exchange tail: totals in rank order => bit-identical on every rank.
__global__ void rank_sum_kernel(const double *gathered, int world, int count, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double v = 0.0;
    for (int r = 0; r < world; ++r) v += gathered[(size_t)r * count + i];
    out[i] = v;
}

// ============================================================================================
// Host side
// ============================================================================================

static thread_local std::string g_err;

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllGather && CommDestroy && GetErrorString;
    }
};
static NcclApi g_nccl;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(CGG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct cgg_handle {
    cgg_config cfg;
    Dev d;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *X_owned = nullptr, *y_owned = nullptr;
    double *replay_dev = nullptr; uint64_t replay_cap = 0;
    double *samples_dev = nullptr; size_t samples_cap = 0;
    double *scratch_dev = nullptr;  // KMAX cand + KMAX out + 1
    double *colstat_dev = nullptr;  // [p][CS_STRIDE]
    size_t smem = 0;
    unsigned long long *prof_dev = nullptr;
    Hdr *hdr_pinned = nullptr;
    bool has_data = false, jet_wanted = false;
    std::vector<char> chain_init, fx_valid, fx_mag;   // fx_mag: the carried f(x0) is at least a valid magnitude (light jet passes)
    int num_sms = 0, max_grid = 0;
    cgg_exchange_fn xfn = nullptr; void *xuser = nullptr;
    ncclComm_t comm = nullptr; int world = 1, rank = 0; double *gather_dev = nullptr; size_t gather_cap = 0;
    double local_ll_const = 0.0;
    MboxEntry *mbox_own = nullptr; size_t mbox_bytes = 0; void *mbox_peer[8] = {nullptr}; bool mbox_ipc[8] = {false}; bool p2p_ready = false;
    uint32_t mbox_stamp = 0;               // passes exchanged through the mailboxes so far (all ranks agree)
    int cluster_S = 0;                    // > 0: the cluster driver runs the sweeps, with this many CTAs per chain
    size_t cluster_smem = 0;
    std::vector<double> w_chain;          // per-chain slice widths (cgg_set_chain_w); default cfg.w
    std::vector<ChainState> last_cs;      // the chains' counters after the last cgg_run (cgg_get_chain_stats)
};

// one component of the prior; returns false (with the message set) if the parameters are invalid
static bool make_prior(int kind, double a, double b, double c, PriorParams &pp) {
    pp.kind = kind; pp.mu = a; pp.sigma = b; pp.df = c; pp.c0 = 0.0; pp.inv_sigma = 1.0;
    if (kind < CGG_PRIOR_NORMAL || kind > CGG_PRIOR_EXPONENTIAL) {
        fail(CGG_E_UNSUPPORTED, "unsupported prior %d; supported: normal, laplace, student_t, gamma, exponential", kind);
        return false;
    }
    if (!(b > 0.0) || !std::isfinite(b)) { fail(CGG_E_ARG, "prior scale / rate must be positive"); return false; }
    if (kind == CGG_PRIOR_STUDENT_T && !(c > 0.0)) { fail(CGG_E_ARG, "student-t df must be positive"); return false; }
    if (kind == CGG_PRIOR_GAMMA && !(a > 0.0)) { fail(CGG_E_ARG, "gamma shape must be positive"); return false; }
    pp.inv_sigma = 1.0 / b;
    if (kind == CGG_PRIOR_NORMAL) pp.c0 = -(kLnSqrt2Pi + log(b));
    else if (kind == CGG_PRIOR_LAPLACE) pp.c0 = -log(2.0 * b);
    else if (kind == CGG_PRIOR_STUDENT_T) pp.c0 = lgamma(0.5 * (c + 1.0)) - lgamma(0.5 * c) - 0.5 * log(c * M_PI) - log(b);
    else if (kind == CGG_PRIOR_GAMMA) pp.c0 = a * log(b) - lgamma(a);
    else pp.c0 = log(b);
    return true;
}

extern "C" const char *cgg_last_error(void) { return g_err.c_str(); }
extern "C" int cgg_abi_version(void) { return CGG_ABI_VERSION; }

static void *kernel_ptr(int family, int which) {     // which: 0 persistent sweep, 1 single pass, 2 cluster sweep
    switch (family * 3 + which) {
    case CGG_GAUSSIAN * 3 + 0: return (void *)sweep_persistent_kernel<CGG_GAUSSIAN>;
    case CGG_GAUSSIAN * 3 + 1: return (void *)pass_kernel<CGG_GAUSSIAN>;
    case CGG_GAUSSIAN * 3 + 2: return (void *)sweep_cluster_kernel<CGG_GAUSSIAN>;
    case CGG_BINOMIAL * 3 + 0: return (void *)sweep_persistent_kernel<CGG_BINOMIAL>;
    case CGG_BINOMIAL * 3 + 1: return (void *)pass_kernel<CGG_BINOMIAL>;
    case CGG_BINOMIAL * 3 + 2: return (void *)sweep_cluster_kernel<CGG_BINOMIAL>;
    case CGG_POISSON * 3 + 0: return (void *)sweep_persistent_kernel<CGG_POISSON>;
    case CGG_POISSON * 3 + 1: return (void *)pass_kernel<CGG_POISSON>;
    case CGG_POISSON * 3 + 2: return (void *)sweep_cluster_kernel<CGG_POISSON>;
    case CGG_KF_NEGBIN * 3 + 1: return (void *)pass_kernel<CGG_KF_NEGBIN>;       // exact passes, stepwise driver only
    case CGG_KF_PROBIT * 3 + 1: return (void *)pass_kernel<CGG_KF_PROBIT>;
    default: return nullptr;
    }
}

// Row-sharded mode: turn the local per-candidate sums in d.xbuf into global sums, identical on every rank.
// Row-sharded mode: turn a device buffer of local sums into global sums, identical on every rank.
static int exchange_buf(cgg_handle *h, double *buf, int64_t count) {
    if (h->comm) {
        if ((size_t)count > h->gather_cap) {
            cudaFree(h->gather_dev); h->gather_dev = nullptr; h->gather_cap = 0;
            CK(cudaMalloc((void **)&h->gather_dev, sizeof(double) * (size_t)h->world * (size_t)count));
            h->gather_cap = (size_t)count;
        }
        ncclResult_t r = g_nccl.AllGather(buf, h->gather_dev, (size_t)count, ncclDouble, h->comm, h->stream);
        if (r != ncclSuccess) return fail(CGG_E_COMM, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
        rank_sum_kernel<<<(int)((count + 127) / 128), 128, 0, h->stream>>>(h->gather_dev, h->world, (int)count, buf);
        CK(cudaGetLastError());
        return CGG_OK;
    }
    if (!h->xfn) return fail(CGG_E_STATE, "row-sharded handle has no exchange (cgg_comm_init_nccl or cgg_set_exchange)");
    if (h->xfn(h->xuser, buf, count, (void *)h->stream) != 0) return fail(CGG_E_COMM, "exchange callback failed");
    return CGG_OK;
}
static int exchange(cgg_handle *h) { return exchange_buf(h, h->d.xbuf, (int64_t)h->d.C * NV); }

static int launch_pass(cgg_handle *h, int mode) {
    Dev d = h->d;
    void *args[] = {&d, &mode};
    CK(cudaLaunchKernel(kernel_ptr(h->cfg.family, 1), dim3(h->d.G), dim3(THREADS), args, h->smem, h->stream));
    return CGG_OK;
}

extern "C" int cgg_create(const cgg_config *cfg, cgg_handle **out) {
    if (!cfg || !out) return fail(CGG_E_ARG, "cgg_create: NULL argument");
    *out = nullptr;
    if (cfg->abi_version != CGG_ABI_VERSION) return fail(CGG_E_ARG, "cgg_create: abi_version %d != %d", cfg->abi_version, CGG_ABI_VERSION);
    if (cfg->n <= 0 || cfg->p <= 0) return fail(CGG_E_ARG, "cgg_create: n and p must be positive");
    if (cfg->n_chains < 1 || cfg->n_chains > CMAX) return fail(CGG_E_ARG, "cgg_create: n_chains must be in 1..%d", CMAX);
    if (cfg->K < 1 || cfg->K > CGG_KMAX) return fail(CGG_E_ARG, "cgg_create: K must be in 1..%d", CGG_KMAX);
    if (!(cfg->w > 0.0) || !std::isfinite(cfg->w)) return fail(CGG_E_ARG, "cgg_create: slice width w must be positive and finite");
    const bool fam_ok = (cfg->family == CGG_GAUSSIAN && cfg->link == CGG_LINK_IDENTITY) ||
                        (cfg->family == CGG_BINOMIAL && (cfg->link == CGG_LINK_LOGIT || cfg->link == CGG_LINK_PROBIT)) ||
                        (cfg->family == CGG_POISSON && cfg->link == CGG_LINK_LOG) ||
                        (cfg->family == CGG_NEGATIVE_BINOMIAL && cfg->link == CGG_LINK_LOG);
    if (!fam_ok) return fail(CGG_E_UNSUPPORTED, "cgg_create: unsupported family/link pair (%d, %d); supported: gaussian/identity, binomial/logit, binomial/probit, poisson/log, negative binomial/log", cfg->family, cfg->link);
    // kernel-side family: binomial + probit is a family of its own; it and the negative binomial run exact passes on the stepwise driver
    const int kf = (cfg->family == CGG_BINOMIAL && cfg->link == CGG_LINK_PROBIT) ? CGG_KF_PROBIT : cfg->family;
    const bool exact_only = kf == CGG_KF_NEGBIN || kf == CGG_KF_PROBIT;
    if (exact_only && cfg->mode == CGG_MODE_ROW_SHARDED) return fail(CGG_E_UNSUPPORTED, "cgg_create: binomial/probit and negative binomial are not available row-sharded");
    PriorParams pp0;
    if (!make_prior(cfg->prior, cfg->prior_mu, cfg->prior_sigma, cfg->prior_df, pp0)) { g_err = "cgg_create: " + g_err; return (cfg->prior < CGG_PRIOR_NORMAL || cfg->prior > CGG_PRIOR_EXPONENTIAL) ? CGG_E_UNSUPPORTED : CGG_E_ARG; }
    if (cfg->family == CGG_GAUSSIAN && !(cfg->sd > 0.0)) return fail(CGG_E_ARG, "cgg_create: gaussian sd must be positive");
    if (cfg->driver != CGG_DRIVER_PERSISTENT && cfg->driver != CGG_DRIVER_STEPWISE && cfg->driver != CGG_DRIVER_CLUSTER) return fail(CGG_E_ARG, "cgg_create: unknown driver");
    if (cfg->mode != CGG_MODE_CHAINS && cfg->mode != CGG_MODE_ROW_SHARDED) return fail(CGG_E_ARG, "cgg_create: unknown mode");
    if (cfg->mode == CGG_MODE_ROW_SHARDED && cfg->driver == CGG_DRIVER_CLUSTER) return fail(CGG_E_ARG, "cgg_create: row-sharded mode runs on the persistent (peer mailboxes) or the stepwise driver");
    if ((cfg->flags & CGG_FLAG_NAIVE) && cfg->mode == CGG_MODE_ROW_SHARDED) return fail(CGG_E_UNSUPPORTED, "cgg_create: the naive linear-predictor mode is not available row-sharded");
    if (cfg->driver == CGG_DRIVER_CLUSTER && cfg->n > (1LL << 22)) return fail(CGG_E_ARG, "cgg_create: the cluster driver is for small n (<= 2^22 rows per chain)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(CGG_E_CUDA, "cgg_create: no CUDA device available (this engine has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(CGG_E_ARG, "cgg_create: device %d out of range (%d devices)", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));

    cgg_handle *h = new cgg_handle();
    h->cfg = *cfg;
    h->cfg.family = kf;                                      // from here on: the kernel-side family
    if (exact_only) { h->cfg.driver = CGG_DRIVER_STEPWISE; h->cfg.flags |= CGG_FLAG_NO_JET | CGG_FLAG_NO_CLUSTER; }
    if (h->cfg.flags & CGG_FLAG_NAIVE) { h->cfg.driver = CGG_DRIVER_STEPWISE; h->cfg.flags |= CGG_FLAG_NO_CLUSTER; }   // host-launched GEMV before every pass
    cfg = &h->cfg;
    memset(&h->d, 0, sizeof(Dev));
    Dev &d = h->d;
    const int C = cfg->n_chains;
    d.n = cfg->n; d.p = cfg->p; d.C = C; d.K = cfg->K; d.family = cfg->family; d.n_total = (double)cfg->n;
    d.inv_sd = 1.0 / (cfg->family == CGG_GAUSSIAN ? cfg->sd : 1.0);
    d.w = cfg->w; d.max_steps = cfg->max_steps < 0 ? -1 : cfg->max_steps;
    d.seed = cfg->seed; d.chain_offset = cfg->chain_offset; d.tau = cfg->spec_tau;
    d.sharded = cfg->mode == CGG_MODE_ROW_SHARDED;
    d.coarse = (cfg->family == CGG_BINOMIAL) && !d.sharded && !(cfg->flags & CGG_FLAG_NO_PREFILTER);
    d.coarse_theta = getenv("CGG_COARSE_THETA") ? atof(getenv("CGG_COARSE_THETA")) : 0.4;
    h->jet_wanted = !(cfg->flags & CGG_FLAG_NO_JET);
    d.jet = h->jet_wanted && !d.sharded;     // row-sharded: decided at cgg_set_data (needs the exchange for the column statistics)
    d.jet_bscale = (cfg->jet_bound_scale > 0.0) ? cfg->jet_bound_scale : 1.0;
    {   // pair passes (chains 2k, 2k + 1 share a walk over the rows when they are at the same coordinate): from 4 chains on
        // the pairs' decisions still hide behind the other pairs' passes; CGG_PAIR=0/1 overrides (experiments, tests)
        const char *e = getenv("CGG_PAIR");
        d.pair = e ? atoi(e) : (C >= 4);
    }
    d.jet_light = d.jet && !(cfg->flags & CGG_FLAG_NO_JET_LIGHT);
    d.prior.comp[0] = pp0; d.prior.n = 1;

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    {   // keep freed pool memory cached on the device (see cgg_set_data)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    h->num_sms = prop.multiProcessorCount;
    int occ = 0;
    h->smem = (CtaShared::bytes(C) + 15) / 16 * 16;
    {   // X-column cache of the pair passes (cgg_device.cuh: ColCache): two columns of every warp's rows in shared memory,
        // if they fit next to the rings.  Needs the grid size, i.e. a first guess of G = all SMs but the deciders'.
        const char *e4 = getenv("CGG_COLCACHE");
        const bool want = (e4 ? atoi(e4) != 0 : true) && d.pair && !d.sharded && cfg->driver == CGG_DRIVER_PERSISTENT;
        const int64_t n_tiles = (cfg->n + TILE_ROWS - 1) / TILE_ROWS;
        const int64_t Wg = (int64_t)(prop.multiProcessorCount - 1) * NWARPS;
        const int64_t tpw = (n_tiles + Wg - 1) / Wg;
        const size_t need = (size_t)NWARPS * 2 * (size_t)tpw * 512;
        const size_t room = (size_t)prop.sharedMemPerBlockOptin > h->smem + 2048 ? (size_t)prop.sharedMemPerBlockOptin - h->smem - 2048 : 0;
        d.colcache = (want && Wg > 0 && n_tiles >= Wg && need <= room) ? (int)tpw : 0;
        h->smem += (size_t)NWARPS * 2 * (size_t)d.colcache * 512;
    }
    if (!exact_only) CK(cudaFuncSetAttribute(kernel_ptr(cfg->family, 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    CK(cudaFuncSetAttribute(kernel_ptr(cfg->family, 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel_ptr(cfg->family, exact_only ? 1 : 0), THREADS, h->smem));
    if (occ < 1) { delete h; return fail(CGG_E_CUDA, "cgg_create: sweep kernel does not fit on an SM"); }
    if (occ > 1) occ = 1;
    h->max_grid = occ * h->num_sms;
    // Workers are warps owning 64-row tiles w, w + W, ...; use fewer CTAs when there are not enough tiles
    // to give every warp at least `tmin` of them (rows_per_cta_min / (64 * NWARPS), default 1).
    d.n_tiles = (d.n + TILE_ROWS - 1) / TILE_ROWS;
    int64_t tmin = cfg->rows_per_cta_min > 0 ? cfg->rows_per_cta_min / (TILE_ROWS * NWARPS) : 1;
    if (tmin < 1) tmin = 1;
    int64_t G = (d.n_tiles + NWARPS * tmin - 1) / (NWARPS * tmin);
    const int64_t gmax = h->max_grid - (cfg->driver == CGG_DRIVER_PERSISTENT ? 1 : 0);   // persistent driver: one more CTA hosts the deciders
    if (gmax < 1) { delete h; return fail(CGG_E_CUDA, "cgg_create: the device cannot co-schedule a worker CTA and the decider CTA"); }
    if (G > gmax) G = gmax;
    if (G > 250) G = 250;      // the limb accumulators count arrivals in 8 bits (cgg_device.cuh: LimbAcc)
    if (G < 1) G = 1;
    d.G = (int)G;
    if (d.colcache && (int64_t)d.colcache * (int64_t)d.G * NWARPS < d.n_tiles) d.colcache = 0;    // (fewer CTAs than assumed: a warp's tiles would not fit)
    {   // group passes (chains 4k .. 4k + 3 share a walk): an experiment, compiled in with -DCGG_GROUP_PASSES only;
        // then on from 12 chains (three groups, so that a group's decisions still hide behind other groups' passes); they need
        // pair passes and the X-column cache.  CGG_QUAD=0/1 overrides.
        const char *e5 = getenv("CGG_QUAD");
        d.quad = ((e5 ? atoi(e5) != 0 : C >= 12) && d.pair && d.colcache > 0) ? 1 : 0;
    }
    d.lde = (d.n + 31) / 32 * 32;
    {   // small n: one cluster per chain instead of the grid-wide protocol
        const char *e3 = getenv("CGG_SMALLN");
        const bool allow = e3 ? atoi(e3) != 0 : true;
        const bool want = cfg->driver == CGG_DRIVER_CLUSTER || (cfg->driver == CGG_DRIVER_PERSISTENT && allow && !(cfg->flags & CGG_FLAG_NO_CLUSTER) && !d.sharded && cfg->n <= (1LL << 18));
        if (want) {
            // CTAs per chain: enough that a worker warp owns a handful of tiles, at most 16 (8 is the portable limit)
            const int64_t per = (int64_t)TILE_ROWS * (NWARPS - 1) * 12;
            int S = 1;
            while (S < CLUSTER_MAX && (int64_t)S * per < cfg->n) S *= 2;
            if (getenv("CGG_CLUSTER")) S = std::max(1, std::min(CLUSTER_MAX, atoi(getenv("CGG_CLUSTER"))));
            void *kf = kernel_ptr(cfg->family, 2);
            h->cluster_smem = ClusterShared::bytes();
            cudaError_t ce = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cluster_smem);
            if (ce == cudaSuccess && S > 8) ce = cudaFuncSetAttribute(kf, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            for (; ce == cudaSuccess && S >= 1; S /= 2) {      // the largest cluster the device can schedule
                cudaLaunchConfig_t lc;
                memset(&lc, 0, sizeof lc);
                lc.gridDim = dim3(S); lc.blockDim = dim3(THREADS); lc.dynamicSmemBytes = h->cluster_smem;
                cudaLaunchAttribute at;
                at.id = cudaLaunchAttributeClusterDimension;
                at.val.clusterDim.x = S; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                lc.attrs = &at; lc.numAttrs = 1;
                int ncl = 0;
                if (S == 1 || (cudaOccupancyMaxActiveClusters(&ncl, kf, &lc) == cudaSuccess && ncl > 0)) break;
                cudaGetLastError();
            }
            if (ce != cudaSuccess || S < 1) {
                cudaGetLastError();
                if (cfg->driver == CGG_DRIVER_CLUSTER) { delete h; return fail(CGG_E_CUDA, "cgg_create: the device cannot run the cluster driver"); }
            } else {
                h->cluster_S = S;
                d.pair = 0; d.quad = 0; d.coarse = 0;       // one chain per cluster; the pre-filter's clamp flags live in global memory
            }
        }
    }
    {   // the plain update decided, published and booked straight from the deciding warp's cache (cgg_device.cuh:
        // jet_fast_update): on wherever a chain's decision sits on its critical cycle -- the cluster driver, one to four
        // chains per device, the row-sharded runs (measured: cfg2 +17 %, one chain +17 %, cfg5's shard +6 %, README shape
        // +15 %) -- and off where three or more pairs of chains hide each other's decisions anyway (8 chains: -0.8 %, +1 %:
        // noise).  CGG_EARLY=0/1 overrides (experiments, tests).  Never changes results.
        const char *e6 = getenv("CGG_EARLY");
        d.early = (e6 ? atoi(e6) != 0 : (h->cluster_S > 0 || !(d.pair && C >= 6))) ? 1 : 0;
    }
    {   // rounding allowance of an accumulated moment: one rounding per add along the longest chain of additions a value goes
        // through -- the rows a lane owns, then the warp, CTA and (row-sharded: rank) folds; the grid fold is exact (limb
        // accumulators / fixed-point accumulators).  Twice that, and never less than JET_CROUND.
        const double depth = std::ceil((double)d.n / ((double)d.G * NWARPS * 32.0)) + 5.0 + NWARPS + 64.0;
        d.jet_ce = std::max(JET_CROUND, 2.0 * depth) * JET_EPS;
    }
    h->w_chain.assign(C, cfg->w);

    // stream-ordered pool allocations (the pool keeps freed memory cached, see above): creating and destroying a handle
    // costs no device-wide synchronising cudaMalloc / cudaFree
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    auto A = [&](void **p, size_t bytes) { return cudaMallocAsync(p, bytes, h->stream); };
    if (e == cudaSuccess) e = A((void **)&d.eta, sizeof(double) * (size_t)C * d.lde);
    if (e == cudaSuccess) e = A((void **)&d.beta, sizeof(double) * (size_t)C * d.p);
    if (e == cudaSuccess) e = A((void **)&d.shat, sizeof(double) * (size_t)C * d.p);
    if (e == cudaSuccess) e = A((void **)&d.acc, sizeof(Acc) * (size_t)C * NV);
    if (e == cudaSuccess) e = A((void **)&d.sync, sizeof(ChainSync) * (size_t)C);
    if (e == cudaSuccess) e = A((void **)&d.xbuf, sizeof(double) * (size_t)C * NV);
    if (e == cudaSuccess) e = A((void **)&d.lacc, sizeof(LimbAcc) * (size_t)C * NV);
    if (e == cudaSuccess) e = A((void **)&d.ctl, sizeof(Ctl) * (size_t)C);
    if (e == cudaSuccess) e = A((void **)&d.cs, sizeof(ChainState) * (size_t)C);
    if (e == cudaSuccess) e = A((void **)&d.hdr, sizeof(Hdr));
    if (e == cudaSuccess) e = A((void **)&h->scratch_dev, sizeof(double) * (3 * KMAX + NV + 2));
    if (e == cudaSuccess) e = A((void **)&h->colstat_dev, sizeof(double) * (size_t)d.p * CS_STRIDE);
    h->hdr_pinned = new Hdr();      // (pageable: a pinned allocation per handle costs more than the 16-byte copies it serves)
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) {
        int rc = fail(CGG_E_CUDA, "cgg_create: allocation failed: %s", cudaGetErrorString(e));
        cgg_destroy(h);
        return rc;
    }
    cudaMemsetAsync(d.eta, 0, sizeof(double) * (size_t)C * d.lde, h->stream);
    cudaMemsetAsync(d.cs, 0, sizeof(ChainState) * (size_t)C, h->stream);
    cudaMemsetAsync(d.ctl, 0, sizeof(Ctl) * (size_t)C, h->stream);
    cudaMemsetAsync(d.hdr, 0, sizeof(Hdr), h->stream);
    cudaMemsetAsync(d.acc, 0, sizeof(Acc) * (size_t)C * NV, h->stream);
    cudaMemsetAsync(d.sync, 0, sizeof(ChainSync) * (size_t)C, h->stream);
    std::vector<double> neg((size_t)C * d.p, -1.0);
    cudaMemcpyAsync(d.shat, neg.data(), sizeof(double) * neg.size(), cudaMemcpyHostToDevice, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    h->chain_init.assign(C, 0);
    h->fx_valid.assign(C, 0);
    h->fx_mag.assign(C, 0);
    d.colstat = h->colstat_dev;
    {   // Keep the chains' eta vectors resident in the 126 MB L2 across passes (cfg3: 8 x 8 MB): they are the only operand
        // that is read AND written every pass, and the only one whose traffic grows with the number of chains.  With eta
        // pinned (persisting lines, set-aside carved out of L2) HBM only carries each X column once per column step.
        // CGG_L2_PERSIST=0 turns it off (experiments).  Never changes results.
        const char *e2 = getenv("CGG_L2_PERSIST");
        const bool want = e2 ? atoi(e2) != 0 : true;
        const size_t eta_bytes = sizeof(double) * (size_t)C * d.lde;
        if (want && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, eta_bytes);
            size_t cur = 0;
            cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
            if (cur < carve) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = (void *)d.eta;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(eta_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)cur / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
        }
    }
    if (getenv("CGG_DEBUG_PTRS"))
        fprintf(stderr, "[cgg ptrs] eta %p beta %p shat %p acc %p sync %p xbuf %p lacc %p ctl %p cs %p hdr %p colstat %p\n", (void *)d.eta, (void *)d.beta, (void *)d.shat,
                (void *)d.acc, (void *)d.sync, (void *)d.xbuf, (void *)d.lacc, (void *)d.ctl, (void *)d.cs, (void *)d.hdr, (void *)h->colstat_dev);
    *out = h;
    return CGG_OK;
}

extern "C" void cgg_destroy(cgg_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    Dev &d = h->d;
    if (h->stream) {
        void *pool_owned[] = {d.eta, d.beta, d.shat, d.acc, d.sync, d.xbuf, d.ctl, d.cs, d.hdr, d.lacc, h->scratch_dev, h->colstat_dev,
                              h->replay_dev, h->samples_dev};
        for (void *q : pool_owned) if (q) cudaFreeAsync(q, h->stream);
    }
    cudaFree(h->prof_dev);
    if (h->X_owned) cudaFreeAsync(h->X_owned, h->stream);
    if (h->y_owned) cudaFreeAsync(h->y_owned, h->stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->gather_dev);
    for (int r = 0; r < 8; ++r) if (h->mbox_ipc[r] && h->mbox_peer[r]) cudaIpcCloseMemHandle(h->mbox_peer[r]);
    cudaFree(h->mbox_own);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    delete h->hdr_pinned;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int finish_set_data(cgg_handle *h) {
    Dev &d = h->d;
    // y support check + lgamma sum: partials summed on the host in CTA order
    const int G = 256;
    double *part = nullptr; int *bad = nullptr;
    CK(cudaMalloc((void **)&part, sizeof(double) * G));
    CK(cudaMalloc((void **)&bad, sizeof(int)));
    CK(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    scan_y_kernel<<<G, THREADS, 0, h->stream>>>(d, part, bad);
    std::vector<double> hp(G);
    int hbad = 0;
    cudaError_t e1 = cudaMemcpyAsync(hp.data(), part, sizeof(double) * G, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e3 = cudaStreamSynchronize(h->stream);
    cudaFree(part); cudaFree(bad);
    CK(e1); CK(e2); CK(e3);
    if (hbad) return fail(CGG_E_ARG, "cgg_set_data: %d response value(s) outside the family's support (binomial needs 0/1, poisson non-negative integers, all finite)", hbad);
    long double s = 0.0L;
    for (int i = 0; i < G; ++i) s += hp[i];
    if (d.family == CGG_GAUSSIAN) d.ll_const = -(double)d.n * (kLnSqrt2Pi + log(h->cfg.sd));
    else if (d.family == CGG_POISSON) d.ll_const = -(double)s;
    else d.ll_const = 0.0;
    if (h->jet_wanted) {
        // column statistics of the jet passes; a row-sharded handle reduces them over the ranks (needs its exchange now)
        d.jet = (!d.sharded || h->comm || h->xfn) ? 1 : 0;
        d.jet_light = d.jet && !(h->cfg.flags & CGG_FLAG_NO_JET_LIGHT);
    }
    if (d.jet) {
        const int grid = (int)std::min<int64_t>(d.p, 4 * (int64_t)h->num_sms);
        double *hist = nullptr, *sums = nullptr;
        CK(cudaMalloc((void **)&hist, sizeof(double) * (size_t)d.p * CS_BINS));
        CK(cudaMalloc((void **)&sums, sizeof(double) * (size_t)d.p * 10));
        int rc = CGG_OK;
        cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)d.p * CS_BINS, h->stream);
        if (e == cudaSuccess) { col_max_kernel<<<grid, THREADS, 0, h->stream>>>(d.X, d.n, d.ldx, d.p, hist); e = cudaGetLastError(); }
        if (e == cudaSuccess && d.sharded) rc = exchange_buf(h, hist, d.p * CS_BINS);
        if (e == cudaSuccess && rc == CGG_OK) { col_sums_kernel<<<grid, THREADS, 0, h->stream>>>(d.X, d.y, d.family, d.n, d.ldx, d.p, hist, sums); e = cudaGetLastError(); }
        double n_total[1] = {(double)d.n};
        if (e == cudaSuccess && rc == CGG_OK && d.sharded) {
            rc = exchange_buf(h, sums, d.p * 10);
            // the global row count (for the rounding allowance of the sums) travels the same way
            if (rc == CGG_OK) {
                e = cudaMemcpyAsync(h->scratch_dev, n_total, sizeof(double), cudaMemcpyHostToDevice, h->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
                if (e == cudaSuccess) rc = exchange_buf(h, h->scratch_dev, 1);
                if (e == cudaSuccess && rc == CGG_OK) e = cudaMemcpyAsync(n_total, h->scratch_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            }
        }
        if (e == cudaSuccess && rc == CGG_OK) {
            col_finish_kernel<<<(int)((d.p + 127) / 128), 128, 0, h->stream>>>(hist, sums, d.p, n_total[0], h->colstat_dev);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(hist); cudaFree(sums);
        if (rc != CGG_OK) return rc;
        CK(e);
        d.n_total = n_total[0];
    }
    h->has_data = true;
    std::fill(h->chain_init.begin(), h->chain_init.end(), 0);
    std::fill(h->fx_valid.begin(), h->fx_valid.end(), 0);
    std::fill(h->fx_mag.begin(), h->fx_mag.end(), 0);
    return CGG_OK;
}

extern "C" int cgg_set_data(cgg_handle *h, const double *X_host, int64_t ldx, const double *y_host) {
    if (!h || !X_host || !y_host) return fail(CGG_E_ARG, "cgg_set_data: NULL argument");
    Dev &d = h->d;
    if (ldx < d.n) return fail(CGG_E_ARG, "cgg_set_data: ldx (%lld) < n (%lld)", (long long)ldx, (long long)d.n);
    CK(cudaSetDevice(h->cfg.device));
    const int64_t ldd = (d.n + 31) / 32 * 32;
    // stream-ordered pool allocation: a later handle on this device reuses the memory instead of paying
    // cudaFree/cudaMalloc of several GB per mcmcglm() call
    if (!h->X_owned) CK(cudaMallocAsync((void **)&h->X_owned, sizeof(double) * (size_t)ldd * d.p, h->stream));
    if (!h->y_owned) CK(cudaMallocAsync((void **)&h->y_owned, sizeof(double) * (size_t)ldd, h->stream));
    CK(cudaMemcpy2DAsync(h->X_owned, sizeof(double) * ldd, X_host, sizeof(double) * ldx, sizeof(double) * d.n, d.p, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->y_owned, y_host, sizeof(double) * d.n, cudaMemcpyHostToDevice, h->stream));
    d.X = h->X_owned; d.y = h->y_owned; d.ldx = ldd;
    return finish_set_data(h);
}

extern "C" int cgg_set_data_device(cgg_handle *h, const double *X_dev, int64_t ldx, const double *y_dev) {
    if (!h || !X_dev || !y_dev) return fail(CGG_E_ARG, "cgg_set_data_device: NULL argument");
    Dev &d = h->d;
    if (ldx < d.n || (ldx & 1)) return fail(CGG_E_ARG, "cgg_set_data_device: ldx must be even and >= n");
    if (((uintptr_t)X_dev & 15) || ((uintptr_t)y_dev & 15)) return fail(CGG_E_ARG, "cgg_set_data_device: X and y must be 16-byte aligned");
    CK(cudaSetDevice(h->cfg.device));
    d.X = X_dev; d.y = y_dev; d.ldx = ldx;
    return finish_set_data(h);
}

static int check_chain(cgg_handle *h, int32_t chain, const char *who, bool need_init) {
    if (!h) return fail(CGG_E_ARG, "%s: NULL handle", who);
    if (!h->has_data) return fail(CGG_E_STATE, "%s: call cgg_set_data first", who);
    if (chain < 0 || chain >= h->d.C) return fail(CGG_E_ARG, "%s: chain %d out of range", who, chain);
    if (need_init && !h->chain_init[chain]) return fail(CGG_E_STATE, "%s: chain %d not initialised (cgg_init_chain)", who, chain);
    return CGG_OK;
}

extern "C" int cgg_init_chain(cgg_handle *h, int32_t chain, const double *beta0_host) {
    int rc = check_chain(h, chain, "cgg_init_chain", false);
    if (rc) return rc;
    if (!beta0_host) return fail(CGG_E_ARG, "cgg_init_chain: NULL beta0");
    Dev &d = h->d;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(d.beta + (int64_t)chain * d.p, beta0_host, sizeof(double) * d.p, cudaMemcpyHostToDevice, h->stream));
    int grid = (int)std::min<int64_t>((d.n / 2 + THREADS - 1) / THREADS + 1, 8 * h->num_sms);
    init_eta_kernel<<<grid, THREADS, 0, h->stream>>>(d, chain);
    CK(cudaGetLastError());
    ChainState cs;
    memset(&cs, 0, sizeof cs);
    cs.phase = PH_START;
    CK(cudaMemcpyAsync(d.cs + chain, &cs, sizeof cs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->chain_init[chain] = 1;
    h->fx_valid[chain] = 0; h->fx_mag[chain] = 0;
    return CGG_OK;
}

extern "C" int cgg_set_state(cgg_handle *h, int32_t chain, const double *beta_host, const double *eta_host) {
    int rc = check_chain(h, chain, "cgg_set_state", false);
    if (rc) return rc;
    if (!beta_host || !eta_host) return fail(CGG_E_ARG, "cgg_set_state: NULL argument");
    Dev &d = h->d;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(d.beta + (int64_t)chain * d.p, beta_host, sizeof(double) * d.p, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d.eta + (int64_t)chain * d.lde, eta_host, sizeof(double) * d.n, cudaMemcpyHostToDevice, h->stream));
    ChainState cs;
    memset(&cs, 0, sizeof cs);
    cs.phase = PH_START;
    CK(cudaMemcpyAsync(d.cs + chain, &cs, sizeof cs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->chain_init[chain] = 1;
    h->fx_valid[chain] = 0; h->fx_mag[chain] = 0;
    return CGG_OK;
}

// Scores up to KMAX candidates of column j of one chain against its current state (no state change).
static int eval_chunk(cgg_handle *h, int32_t chain, int64_t j, int K, const double *cand, double *out, double *prior_sum) {
    Dev &d = h->d;
    std::vector<double> beta_j(1);
    CK(cudaMemcpyAsync(beta_j.data(), d.beta + (int64_t)chain * d.p + j, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<Ctl> ctl(d.C);
    memset(ctl.data(), 0, sizeof(Ctl) * d.C);
    for (int c = 0; c < d.C; ++c) ctl[c].commit_j = -1;
    ctl[chain].j = (int32_t)j; ctl[chain].ncand = K;
    for (int k = 0; k < K; ++k) ctl[chain].delta[k] = cand[k] - beta_j[0];  // diff_beta, R/glm_utils.R:127
    std::vector<Ctl> saved(d.C);
    CK(cudaMemcpyAsync(saved.data(), d.ctl, sizeof(Ctl) * d.C, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(d.ctl, ctl.data(), sizeof(Ctl) * d.C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->scratch_dev, cand, sizeof(double) * K, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(&d.hdr->done, 0, sizeof(int32_t), h->stream));
    int rc = launch_pass(h, 1);
    if (rc) return rc;
    if (h->d.sharded) {
        rc = exchange(h);
        if (rc) return rc;
    }
    finalize_eval_kernel<<<1, 32, 0, h->stream>>>(d, chain, (int)j, K, h->scratch_dev, h->scratch_dev + KMAX, h->scratch_dev + 2 * KMAX);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->scratch_dev + KMAX, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
    if (prior_sum) CK(cudaMemcpyAsync(prior_sum, h->scratch_dev + 2 * KMAX, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(d.ctl, saved.data(), sizeof(Ctl) * d.C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return CGG_OK;
}

extern "C" int cgg_log_potential(cgg_handle *h, int32_t chain, int64_t j, int32_t K, const double *cand_host, double *out_host) {
    int rc = check_chain(h, chain, "cgg_log_potential", true);
    if (rc) return rc;
    if (j < 0 || j >= h->d.p) return fail(CGG_E_ARG, "cgg_log_potential: j out of range");
    if (K < 0 || (K > 0 && (!cand_host || !out_host))) return fail(CGG_E_ARG, "cgg_log_potential: bad candidate buffer");
    CK(cudaSetDevice(h->cfg.device));
    for (int k0 = 0; k0 < K; k0 += KMAX) {
        const int kk = std::min(KMAX, K - k0);
        rc = eval_chunk(h, chain, j, kk, cand_host + k0, out_host + k0, nullptr);
        if (rc) return rc;
    }
    return CGG_OK;
}

// f at the chain's current point (the value qslice's first f(x) would return) + prior sum
static int ensure_fx(cgg_handle *h, int32_t chain) {
    if (h->fx_valid[chain]) return CGG_OK;
    Dev &d = h->d;
    double b0 = 0.0, fx = 0.0, ps = 0.0;
    CK(cudaMemcpyAsync(&b0, d.beta + (int64_t)chain * d.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int rc = eval_chunk(h, chain, 0, 1, &b0, &fx, &ps);
    if (rc) return rc;
    if (fx != fx) return fail(CGG_E_NAN, "log-potential at the starting point of chain %d is NaN", chain);
    ChainState cs;
    CK(cudaMemcpyAsync(&cs, d.cs + chain, sizeof cs, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cs.fx0 = fx; cs.prior_sum = ps;
    CK(cudaMemcpyAsync(d.cs + chain, &cs, sizeof cs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->fx_valid[chain] = 1;
    return CGG_OK;
}

extern "C" int cgg_get_fx(cgg_handle *h, int32_t chain, double *fx) {
    int rc = check_chain(h, chain, "cgg_get_fx", true);
    if (rc) return rc;
    if (!fx) return fail(CGG_E_ARG, "cgg_get_fx: NULL output");
    CK(cudaSetDevice(h->cfg.device));
    rc = ensure_fx(h, chain);
    if (rc) return rc;
    ChainState cs;
    CK(cudaMemcpy(&cs, h->d.cs + chain, sizeof cs, cudaMemcpyDeviceToHost));
    *fx = cs.fx0;
    return CGG_OK;
}

extern "C" int cgg_update_eta(cgg_handle *h, int32_t chain, int64_t j, double new_beta_j) {
    int rc = check_chain(h, chain, "cgg_update_eta", true);
    if (rc) return rc;
    Dev &d = h->d;
    if (j < 0 || j >= d.p) return fail(CGG_E_ARG, "cgg_update_eta: j out of range");
    CK(cudaSetDevice(h->cfg.device));
    double cur = 0.0;
    CK(cudaMemcpyAsync(&cur, d.beta + (int64_t)chain * d.p + j, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double diff = new_beta_j - cur;
    int grid = (int)std::min<int64_t>((d.n / 2 + THREADS - 1) / THREADS + 1, 8 * h->num_sms);
    axpy_eta_kernel<<<grid, THREADS, 0, h->stream>>>(d, chain, j, diff);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(d.beta + (int64_t)chain * d.p + j, &new_beta_j, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->fx_valid[chain] = 0; h->fx_mag[chain] = 0;
    return CGG_OK;
}

extern "C" int cgg_get_state(cgg_handle *h, int32_t chain, double *beta_host, double *eta_host) {
    int rc = check_chain(h, chain, "cgg_get_state", true);
    if (rc) return rc;
    Dev &d = h->d;
    CK(cudaSetDevice(h->cfg.device));
    if (beta_host) CK(cudaMemcpyAsync(beta_host, d.beta + (int64_t)chain * d.p, sizeof(double) * d.p, cudaMemcpyDeviceToHost, h->stream));
    if (eta_host) CK(cudaMemcpyAsync(eta_host, d.eta + (int64_t)chain * d.lde, sizeof(double) * d.n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return CGG_OK;
}

extern "C" int cgg_debug_row_terms(int32_t device, int32_t family, int64_t n, const double *y_host, const double *eta_host,
                                   double sd, double *out_host) {
    if (n <= 0 || !y_host || !eta_host || !out_host) return fail(CGG_E_ARG, "cgg_debug_row_terms: bad argument");
    if (family < CGG_GAUSSIAN || family > CGG_KF_PROBIT) return fail(CGG_E_UNSUPPORTED, "cgg_debug_row_terms: unsupported family (0 gaussian, 1 binomial/logit, 2 poisson, 3 negative binomial, 4 binomial/probit)");
    CK(cudaSetDevice(device));
    double *dy = nullptr, *de = nullptr, *dout = nullptr;
    CK(cudaMalloc((void **)&dy, sizeof(double) * n));
    CK(cudaMalloc((void **)&de, sizeof(double) * n));
    CK(cudaMalloc((void **)&dout, sizeof(double) * n));
    cudaError_t e = cudaMemcpy(dy, y_host, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(de, eta_host, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { row_terms_kernel<<<64, 256>>>(family, n, dy, de, 1.0 / sd, dout); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(out_host, dout, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaFree(dy); cudaFree(de); cudaFree(dout);
    CK(e);
    return CGG_OK;
}

extern "C" int cgg_debug_coarse_error(int32_t device, double *max_err_over_1_plus_abs_s, double *at_s) {
    if (!max_err_over_1_plus_abs_s || !at_s) return fail(CGG_E_ARG, "cgg_debug_coarse_error: NULL output");
    CK(cudaSetDevice(device));
    const int G = 1024;
    double *dv = nullptr; unsigned long long *db = nullptr;
    CK(cudaMalloc((void **)&dv, sizeof(double) * G));
    CK(cudaMalloc((void **)&db, sizeof(unsigned long long) * G));
    coarse_error_scan_kernel<<<G, 256>>>(dv, db);
    std::vector<double> hv(G); std::vector<unsigned long long> hb(G);
    cudaError_t e = cudaMemcpy(hv.data(), dv, sizeof(double) * G, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(hb.data(), db, sizeof(unsigned long long) * G, cudaMemcpyDeviceToHost);
    cudaFree(dv); cudaFree(db);
    CK(e);
    double w = 0.0; unsigned int bits = 0;
    for (int i = 0; i < G; ++i) if (hv[i] > w) { w = hv[i]; bits = (unsigned int)hb[i]; }
    float f; memcpy(&f, &bits, 4);
    *max_err_over_1_plus_abs_s = w; *at_s = (double)f;
    return CGG_OK;
}

extern "C" int cgg_debug_light_error(int32_t device, double *max_abs_err) {
    if (!max_abs_err) return fail(CGG_E_ARG, "cgg_debug_light_error: NULL output");
    CK(cudaSetDevice(device));
    const int G = 1024;
    double *dv = nullptr;
    CK(cudaMalloc((void **)&dv, sizeof(double) * G));
    light_error_scan_kernel<<<G, 256>>>(dv, 256);            // 6.7e7 grid points
    std::vector<double> hv(G);
    cudaError_t e = cudaMemcpy(hv.data(), dv, sizeof(double) * G, cudaMemcpyDeviceToHost);
    cudaFree(dv);
    CK(e);
    double w = 0.0;
    for (int i = 0; i < G; ++i) w = (hv[i] > w || hv[i] != hv[i]) ? hv[i] : w;
    *max_abs_err = w;
    return CGG_OK;
}

extern "C" int cgg_debug_jet(cgg_handle *h, int32_t chain, int64_t j, int32_t K, int32_t light, const double *cand_host,
                             double *value_host, double *bound_host, double *sums_host) {
    int rc = check_chain(h, chain, "cgg_debug_jet", true);
    if (rc) return rc;
    Dev &d = h->d;
    if (!d.jet) return fail(CGG_E_STATE, "cgg_debug_jet: handle was created without jet passes");
    if (j < 0 || j >= d.p) return fail(CGG_E_ARG, "cgg_debug_jet: j out of range");
    if (K < 1 || K > KMAX || !cand_host || !value_host || !bound_host) return fail(CGG_E_ARG, "cgg_debug_jet: K must be in 1..%d and buffers non-NULL", KMAX);
    CK(cudaSetDevice(h->cfg.device));
    std::vector<Ctl> ctl(d.C), saved(d.C);
    memset(ctl.data(), 0, sizeof(Ctl) * d.C);
    for (int c = 0; c < d.C; ++c) ctl[c].commit_j = -1;
    const bool lt = light && d.family == CGG_BINOMIAL;
    double fmag_light = 1.0;
    if (lt) {      // the magnitude a chain would carry: |f(x0)|
        rc = ensure_fx(h, chain);
        if (rc) return rc;
        ChainState cs1;
        CK(cudaMemcpy(&cs1, d.cs + chain, sizeof cs1, cudaMemcpyDeviceToHost));
        fmag_light = fabs(cs1.fx0) + 1.0;
    }
    ctl[chain].j = (int32_t)j; ctl[chain].ncand = 0; ctl[chain].coarse_mask = (int32_t)(JET_BIT | (lt ? 0u : JET_FULL));
    CK(cudaMemcpyAsync(&ctl[chain].cscale, h->colstat_dev + j * CS_STRIDE, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyAsync(saved.data(), d.ctl, sizeof(Ctl) * d.C, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(d.ctl, ctl.data(), sizeof(Ctl) * d.C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->scratch_dev, cand_host, sizeof(double) * K, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(&d.hdr->done, 0, sizeof(int32_t), h->stream));
    rc = launch_pass(h, 1);
    if (rc) return rc;
    jet_debug_kernel<<<1, 32, 0, h->stream>>>(d, chain, (int)j, K, lt ? 1 : 0, fmag_light, h->scratch_dev, h->scratch_dev + KMAX);
    CK(cudaGetLastError());
    std::vector<double> out(2 * K + NV);
    CK(cudaMemcpyAsync(out.data(), h->scratch_dev + KMAX, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(d.ctl, saved.data(), sizeof(Ctl) * d.C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < K; ++k) { value_host[k] = out[k]; bound_host[k] = out[K + k]; }
    if (sums_host) for (int k = 0; k < NV; ++k) sums_host[k] = out[2 * K + k];
    return CGG_OK;
}

extern "C" int cgg_set_exchange(cgg_handle *h, cgg_exchange_fn fn, void *user) {
    if (!h) return fail(CGG_E_ARG, "cgg_set_exchange: NULL handle");
    h->xfn = fn; h->xuser = user;
    return CGG_OK;
}

extern "C" int cgg_nccl_unique_id(char out[CGG_NCCL_ID_BYTES]) {
    if (!out) return fail(CGG_E_ARG, "cgg_nccl_unique_id: NULL output");
    if (!g_nccl.load()) return fail(CGG_E_COMM, "cannot load NCCL (libnccl.so.2): %s", dlerror());
    static_assert(sizeof(ncclUniqueId) == CGG_NCCL_ID_BYTES, "NCCL unique id size");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(CGG_E_COMM, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
    memcpy(out, &id, sizeof id);
    return CGG_OK;
}

extern "C" int cgg_comm_init_nccl(cgg_handle *h, int32_t rank, int32_t world, const char id_bytes[CGG_NCCL_ID_BYTES]) {
    if (!h || !id_bytes) return fail(CGG_E_ARG, "cgg_comm_init_nccl: NULL argument");
    if (!h->d.sharded) return fail(CGG_E_STATE, "cgg_comm_init_nccl: handle was not created with CGG_MODE_ROW_SHARDED");
    if (world < 1 || rank < 0 || rank >= world) return fail(CGG_E_ARG, "cgg_comm_init_nccl: bad rank/world");
    if (!g_nccl.load()) return fail(CGG_E_COMM, "cannot load NCCL (libnccl.so.2): %s", dlerror());
    CK(cudaSetDevice(h->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) { h->comm = nullptr; return fail(CGG_E_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    h->world = world; h->rank = rank;
    CK(cudaMalloc((void **)&h->gather_dev, sizeof(double) * (size_t)world * h->d.C * NV));
    h->gather_cap = (size_t)h->d.C * NV;
    return CGG_OK;
}

extern "C" int cgg_p2p_mailbox(cgg_handle *h, int32_t world, void **dev_ptr, char ipc_handle[CGG_IPC_HANDLE_BYTES]) {
    if (!h || world < 1 || world > 8) return fail(CGG_E_ARG, "cgg_p2p_mailbox: world must be in 1..8");
    if (!h->d.sharded) return fail(CGG_E_STATE, "cgg_p2p_mailbox: handle was not created with CGG_MODE_ROW_SHARDED");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->mbox_own) {
        h->mbox_bytes = sizeof(MboxEntry) * (size_t)h->d.C * (size_t)world * NV;
        CK(cudaMalloc((void **)&h->mbox_own, h->mbox_bytes));       // (not from the pool: must be exportable)
        CK(cudaMemset(h->mbox_own, 0, h->mbox_bytes));              // stamp 0 = never written
        CK(cudaDeviceSynchronize());
    }
    if (dev_ptr) *dev_ptr = h->mbox_own;
    if (ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == CGG_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t ih;
        CK(cudaIpcGetMemHandle(&ih, h->mbox_own));
        memcpy(ipc_handle, &ih, sizeof ih);
    }
    h->world = world;
    return CGG_OK;
}

extern "C" int cgg_p2p_connect(cgg_handle *h, int32_t rank, int32_t world, void *const *dev_ptrs, const char *ipc_handles) {
    if (!h || (!dev_ptrs && !ipc_handles)) return fail(CGG_E_ARG, "cgg_p2p_connect: NULL argument");
    if (!h->mbox_own || world != h->world || rank < 0 || rank >= world) return fail(CGG_E_STATE, "cgg_p2p_connect: call cgg_p2p_mailbox(world) first, with the same world");
    CK(cudaSetDevice(h->cfg.device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->mbox_peer[r] = h->mbox_own; continue; }
        if (dev_ptrs && dev_ptrs[r]) { h->mbox_peer[r] = dev_ptrs[r]; continue; }         // same process (tests: two shards on one device)
        if (!ipc_handles) return fail(CGG_E_ARG, "cgg_p2p_connect: no pointer and no IPC handle for rank %d", r);
        cudaIpcMemHandle_t ih;
        memcpy(&ih, ipc_handles + (size_t)r * CGG_IPC_HANDLE_BYTES, sizeof ih);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(CGG_E_COMM, "cgg_p2p_connect: cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
        h->mbox_peer[r] = p; h->mbox_ipc[r] = true;
    }
    h->rank = rank;
    h->p2p_ready = true;
    return CGG_OK;
}

extern "C" int cgg_add_prior(cgg_handle *h, int32_t kind, double a, double b, double c) {
    if (!h) return fail(CGG_E_ARG, "cgg_add_prior: NULL handle");
    if (h->d.prior.n >= CGG_MAX_PRIORS) return fail(CGG_E_UNSUPPORTED, "cgg_add_prior: at most %d priors in a list", CGG_MAX_PRIORS);
    PriorParams pp;
    if (!make_prior(kind, a, b, c, pp)) { g_err = "cgg_add_prior: " + g_err; return (kind < CGG_PRIOR_NORMAL || kind > CGG_PRIOR_EXPONENTIAL) ? CGG_E_UNSUPPORTED : CGG_E_ARG; }
    h->d.prior.comp[h->d.prior.n++] = pp;
    std::fill(h->fx_valid.begin(), h->fx_valid.end(), 0);
    std::fill(h->fx_mag.begin(), h->fx_mag.end(), 0);
    return CGG_OK;
}

extern "C" int cgg_set_chain_w(cgg_handle *h, const double *w_host) {
    if (!h || !w_host) return fail(CGG_E_ARG, "cgg_set_chain_w: NULL argument");
    for (int c = 0; c < h->d.C; ++c)
        if (!(w_host[c] > 0.0) || !std::isfinite(w_host[c])) return fail(CGG_E_ARG, "cgg_set_chain_w: w[%d] must be positive and finite", c);
    h->w_chain.assign(w_host, w_host + h->d.C);
    return CGG_OK;
}

extern "C" int cgg_get_chain_stats(cgg_handle *h, int32_t chain, cgg_stats *out) {
    if (!h || !out) return fail(CGG_E_ARG, "cgg_get_chain_stats: NULL argument");
    if (chain < 0 || chain >= h->d.C) return fail(CGG_E_ARG, "cgg_get_chain_stats: chain %d out of range", chain);
    if ((int)h->last_cs.size() != h->d.C) return fail(CGG_E_STATE, "cgg_get_chain_stats: no cgg_run yet");
    const ChainState &cs = h->last_cs[chain];
    memset(out, 0, sizeof *out);
    out->updates = cs.updates; out->passes = cs.passes; out->chain_passes = cs.chain_passes; out->commit_passes = cs.commit_passes;
    out->cand_evals = cs.cand_evals; out->ref_evals = cs.ref_evals; out->stepouts = cs.stepouts; out->shrinks = cs.shrinks;
    out->coarse_evals = cs.coarse_evals; out->coarse_undecided = cs.coarse_undecided;
    out->jet_passes = cs.jet_passes; out->jet_fallbacks = cs.jet_fallbacks; out->jet_retries = cs.jet_retries;
    out->algorithmic_bytes = 8.0 * (double)h->d.n * (3.0 * (double)cs.chain_passes + 2.0 * (double)cs.commit_passes);
    return CGG_OK;
}

extern "C" void *cgg_stream(cgg_handle *h) { return h ? (void *)h->stream : nullptr; }

extern "C" int cgg_launch_shape(cgg_handle *h, int32_t *ctas, int32_t *threads) {
    if (!h) return fail(CGG_E_ARG, "cgg_launch_shape: NULL handle");
    if (ctas) *ctas = h->cluster_S > 0 ? h->d.C * h->cluster_S : h->d.G + (h->cfg.driver == CGG_DRIVER_PERSISTENT ? 1 : 0);   // + the decider CTA
    if (threads) *threads = THREADS;
    return CGG_OK;
}

extern "C" int cgg_run(cgg_handle *h, int64_t n_iter, const double *replay_u, uint64_t n_u, uint64_t *u_consumed,
                       double *samples_out, cgg_stats *stats) {
    if (!h) return fail(CGG_E_ARG, "cgg_run: NULL handle");
    if (!h->has_data) return fail(CGG_E_STATE, "cgg_run: call cgg_set_data first");
    if (n_iter < 0) return fail(CGG_E_ARG, "cgg_run: n_iter must be >= 0");
    Dev &d = h->d;
    const int C = d.C;
    for (int c = 0; c < C; ++c)
        if (!h->chain_init[c]) return fail(CGG_E_STATE, "cgg_run: chain %d not initialised (cgg_init_chain)", c);
    const bool mailboxes = d.sharded && h->cfg.driver == CGG_DRIVER_PERSISTENT;
    if (mailboxes && !h->p2p_ready) return fail(CGG_E_STATE, "cgg_run: a row-sharded handle on the persistent driver exchanges through peer mailboxes (cgg_p2p_mailbox + cgg_p2p_connect)");
    if (d.sharded && !mailboxes && !h->xfn && !h->comm) return fail(CGG_E_STATE, "cgg_run: row-sharded handle has no exchange (cgg_comm_init_nccl or cgg_set_exchange)");
    CK(cudaSetDevice(h->cfg.device));
    const bool light = d.jet_light && d.family == CGG_BINOMIAL;
    for (int c = 0; c < C; ++c) {
        if (light && h->fx_mag[c]) continue;     // light passes only use |f(x0)| as a magnitude: the carried value will do
        int rc = ensure_fx(h, c);
        if (rc) return rc;
        h->fx_mag[c] = 1;
    }
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_iter == 0) return CGG_OK;

    // replay stream
    if (replay_u) {
        const uint64_t need = (uint64_t)C * n_u;
        if (need > h->replay_cap) {
            if (h->replay_dev) cudaFreeAsync(h->replay_dev, h->stream);
            h->replay_dev = nullptr; h->replay_cap = 0;
            CK(cudaMallocAsync((void **)&h->replay_dev, sizeof(double) * (need ? need : 1), h->stream));
            h->replay_cap = need;
        }
        CK(cudaMemcpyAsync(h->replay_dev, replay_u, sizeof(double) * need, cudaMemcpyHostToDevice, h->stream));
        d.replay = h->replay_dev; d.n_u = n_u;
    } else { d.replay = nullptr; d.n_u = 0; }
    // sample store
    const size_t ns = (size_t)C * n_iter * d.p;
    if (ns > h->samples_cap) {
        if (h->samples_dev) cudaFreeAsync(h->samples_dev, h->stream);
        h->samples_dev = nullptr; h->samples_cap = 0;
        CK(cudaMallocAsync((void **)&h->samples_dev, sizeof(double) * ns, h->stream));
        h->samples_cap = ns;
    }
    d.samples = h->samples_dev;

    // per-run reset of the chain machines (cursor persists across runs; counters restart)
    std::vector<ChainState> cs(C);
    CK(cudaMemcpyAsync(cs.data(), d.cs, sizeof(ChainState) * C, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<Ctl> ctl(C);
    memset(ctl.data(), 0, sizeof(Ctl) * C);
    for (int c = 0; c < C; ++c) {
        cs[c].phase = PH_START; cs[c].status = CGG_OK; cs[c].iter = 0; cs[c].j = 0;
        cs[c].updates = cs[c].chain_passes = cs[c].commit_passes = cs[c].cand_evals = 0;
        cs[c].ref_evals = cs[c].stepouts = cs[c].shrinks = cs[c].passes = cs[c].coarse_evals = cs[c].coarse_undecided = 0;
        cs[c].jet_passes = cs[c].jet_fallbacks = cs[c].jet_retries = 0;
        cs[c].fine_next = 0; cs[c].jet_skip = 0;
        cs[c].w = h->w_chain[c];
        d.replay_origin[c] = cs[c].cursor;
        ctl[c].commit_j = -1;
    }
    Hdr hdr;
    memset(&hdr, 0, sizeof hdr);
    d.n_iter = n_iter;
    CK(cudaMemcpyAsync(d.cs, cs.data(), sizeof(ChainState) * C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d.ctl, ctl.data(), sizeof(Ctl) * C, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d.hdr, &hdr, sizeof hdr, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(d.acc, 0, sizeof(Acc) * (size_t)C * NV, h->stream));
    CK(cudaMemsetAsync(d.sync, 0, sizeof(ChainSync) * (size_t)C, h->stream));
    CK(cudaMemsetAsync(d.lacc, 0, sizeof(LimbAcc) * (size_t)C * NV, h->stream));   // counts restart with the versions
#ifdef CGG_PROFILE_BUILD
    const bool want_prof = getenv("CGG_PROFILE") != nullptr;
#else
    const bool want_prof = false;
    if (getenv("CGG_PROFILE")) {
        static bool told = false;
        if (!told) fprintf(stderr, "[cgg profile] this libcggibbs.so has no phase counters: rebuild with CGG_NVCC_EXTRA=-DCGG_PROFILE_BUILD python -m mcmcglm_b200.build -f\n");
        told = true;
    }
#endif
    if (want_prof) {
        if (!h->prof_dev) CK(cudaMalloc((void **)&h->prof_dev, 8 * (32 + 4 * 1024 + 32 * 128 * 4 + 2 * 1024 * 32 + 32 * 128)));
        CK(cudaMemsetAsync(h->prof_dev, 0, 8 * (32 + 4 * 1024 + 32 * 128 * 4 + 2 * 1024 * 32 + 32 * 128), h->stream));
        {   // arrival minima start at +inf
            std::vector<unsigned long long> init((size_t)32 * 128 * 4, 0ULL);
            for (size_t i = 2; i < init.size(); i += 4) init[i] = ~0ULL;
            CK(cudaMemcpyAsync(h->prof_dev + 32 + 4096, init.data(), 8 * init.size(), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
    }
    d.prof = want_prof ? h->prof_dev : nullptr;

    uint64_t launches = 0;
    CK(cudaEventRecord(h->ev0, h->stream));
    d.iter_stop = n_iter;
    if (h->cluster_S > 0) {
        Dev dd = d;
        dd.G = h->cluster_S;
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3(C * h->cluster_S); lc.blockDim = dim3(THREADS); lc.dynamicSmemBytes = h->cluster_smem; lc.stream = h->stream;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = h->cluster_S; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        lc.attrs = &at; lc.numAttrs = 1;
        void *args[] = {&dd};
        CK(cudaLaunchKernelExC(&lc, kernel_ptr(h->cfg.family, 2), args));
        ++launches;
    } else if (h->cfg.driver == CGG_DRIVER_PERSISTENT) {
        // With pair passes a long run is cut into launches of a few iterations: every launch starts all chains at
        // coordinate 0 of the same iteration, which brings the chains of a pair back together after an exact-pass
        // hand-over has cost one of them extra passes (inside a launch nothing re-aligns them).  A launch boundary costs
        // ~0.1 ms: the pending eta updates are flushed, the chains re-armed on the device, no host round trip.
        const int64_t chunk = d.pair ? (getenv("CGG_CHUNK") ? std::max(1, atoi(getenv("CGG_CHUNK"))) : 8) : n_iter;
        for (int64_t done = 0; done < n_iter; done += chunk) {
            if (done > 0) {
                rearm_kernel<<<1, 32, 0, h->stream>>>(d);
                CK(cudaGetLastError());
                CK(cudaMemsetAsync(d.lacc, 0, sizeof(LimbAcc) * (size_t)C * NV, h->stream));
            }
            Dev dd = d;
            dd.iter_stop = std::min(done + chunk, n_iter);
            if (mailboxes) {
                for (int r = 0; r < 8; ++r) dd.mbox[r] = (unsigned long long *)h->mbox_peer[r];
                dd.world = h->world; dd.rank = h->rank; dd.mbox_stamp0 = h->mbox_stamp;
            }
            void *args[] = {&dd};
            CK(cudaLaunchCooperativeKernel(kernel_ptr(h->cfg.family, 0), dim3(d.G + 1), dim3(THREADS), args, h->smem, h->stream));
            ++launches;
        }
    } else {
        const int batch = (d.sharded && !h->comm) ? 1 : 32;   // host callbacks are synchronous; NCCL and kernels queue up
        for (;;) {
            for (int i = 0; i < batch; ++i) {
                if (h->cfg.flags & CGG_FLAG_NAIVE) {        // eta <- X %*% beta for every chain, O(n p), before every pass
                    const int grid = (int)std::min<int64_t>((d.n / 2 + THREADS - 1) / THREADS + 1, 8 * h->num_sms);
                    for (int c = 0; c < C; ++c) init_eta_kernel<<<grid, THREADS, 0, h->stream>>>(d, c);
                    naive_drop_commit_kernel<<<(C + 31) / 32, 32, 0, h->stream>>>(d);
                    launches += C + 1;
                }
                int rc = launch_pass(h, d.sharded ? 1 : 0);
                if (rc) return rc;
                ++launches;
                if (d.sharded) {
                    Dev dd = d;
                    if (h->comm) {      // all-gather only: the decider adds the ranks' parts itself, in rank order
                        ncclResult_t r = g_nccl.AllGather(d.xbuf, h->gather_dev, (size_t)d.C * NV, ncclDouble, h->comm, h->stream);
                        if (r != ncclSuccess) return fail(CGG_E_COMM, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
                        dd.gathered = h->gather_dev; dd.world = h->world;
                    } else {
                        rc = exchange(h);
                        if (rc) return rc;
                    }
                    decide_kernel<<<1, THREADS, 0, h->stream>>>(dd);
                    launches += h->comm ? 2 : 1;
                }
            }
            CK(cudaMemcpyAsync(h->hdr_pinned, d.hdr, sizeof(Hdr), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (h->hdr_pinned->done) break;
        }
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (want_prof) {
        std::vector<double> sh((size_t)C * d.p), bt((size_t)C * d.p);
        CK(cudaMemcpy(sh.data(), d.shat, sizeof(double) * sh.size(), cudaMemcpyDeviceToHost));
        double m = 0; int cnt = 0; double mn = 1e300, mx = 0;
        for (double v : sh) if (v > 0) { m += v; ++cnt; mn = std::min(mn, v); mx = std::max(mx, v); }
        fprintf(stderr, "[cgg profile] slice-width estimate shat: mean %.4g min %.4g max %.4g (%d set)\n", cnt ? m / cnt : 0.0, mn, mx, cnt);
    }
    if (want_prof && h->cluster_S > 0) {
        unsigned long long pr[32];
        CK(cudaMemcpy(pr, h->prof_dev, sizeof pr, cudaMemcpyDeviceToHost));
        const double np_ = pr[24] ? (double)pr[24] : 1.0;
        fprintf(stderr, "[cgg profile] cluster driver, %d CTAs per chain, %.3f ms, %llu passes; mean cycles per pass (CTA 0): decider's prefetch + pre-phase %.0f | "
                        "a worker warp's pass %.0f | cluster barrier (decider) %.0f | sums + decision + control block %.0f\n",
                h->cluster_S, ms, pr[24], pr[20] / np_, pr[22] / np_, pr[23] / np_, pr[21] / np_);
        if (pr[12] | pr[14] | pr[15])      // (-DCGG_DECIDER_TICKS builds)
            fprintf(stderr, "[cgg profile] decision phases, mean cycles: state + sums %.0f | draws / points at hand %.0f | round 1 judged %.0f | stepping out + proposals %.0f | accepted %.0f | next pass + store %.0f | fence %.0f\n",
                    pr[12] / np_, pr[14] / np_, pr[15] / np_, pr[16] / np_, pr[17] / np_, pr[18] / np_, pr[19] / np_);
    }
    if (want_prof && h->cfg.driver == CGG_DRIVER_PERSISTENT && h->cluster_S == 0) {
        unsigned long long pr[24];
        CK(cudaMemcpy(pr, h->prof_dev, sizeof pr, cudaMemcpyDeviceToHost));
        const double nw = pr[6] ? (double)pr[6] : 1.0;
        fprintf(stderr, "[cgg profile] %.3f ms; per-worker mean cycles: wait %.3g rows %.3g (tile loop %.3g) arrive %.3g | slow-waits/worker %.1f prefetched-or-pair-passes/worker %.1f workers %llu\n",
                ms, pr[0] / nw, pr[1] / nw, pr[3] / nw, pr[2] / nw, pr[5] / nw, pr[4] / nw, pr[6]);
        fprintf(stderr, "[cgg profile] decisions %llu, mean cycles per decision %.0f | polls that found the decision not yet published %llu, look-aheads %llu (published: %llu)\n",
                pr[8], pr[8] ? (double)pr[7] / (double)pr[8] : 0.0, pr[9], pr[10], pr[11]);
        fprintf(stderr, "[cgg profile] decision phases, mean cycles: load+sums %.0f | uniforms %.0f | step-out %.0f | proposals %.0f | judge+accept %.0f | next pass+store %.0f | fence %.0f\n",
                pr[12] / (double)(pr[8] ? pr[8] : 1), pr[14] / (double)(pr[8] ? pr[8] : 1), pr[15] / (double)(pr[8] ? pr[8] : 1), pr[16] / (double)(pr[8] ? pr[8] : 1),
                pr[17] / (double)(pr[8] ? pr[8] : 1), pr[18] / (double)(pr[8] ? pr[8] : 1), pr[19] / (double)(pr[8] ? pr[8] : 1));
        if (getenv("CGG_PROFILE_TRACE")) {
            std::vector<unsigned long long> tr((size_t)d.C * 128 * 4);
            CK(cudaMemcpy(tr.data(), h->prof_dev + 32 + 4096, 8 * tr.size(), cudaMemcpyDeviceToHost));
            const unsigned long long t0 = tr[(0 * 128 + 0) * 4 + 0];
            std::vector<unsigned long long> need((size_t)32 * 128);
            CK(cudaMemcpy(need.data(), h->prof_dev + 32 + 4096 + 32 * 128 * 4 + 2 * 1024 * 32, 8 * need.size(), cudaMemcpyDeviceToHost));
            for (int v = 40; v < 44; ++v)
                for (int c = 0; c < d.C; ++c) {
                    const unsigned long long *q = &tr[((size_t)c * 128 + v) * 4];
                    fprintf(stderr, "[cgg trace] pass %d chain %d: first warp done %.1f us, last warp done %.1f us, decider saw it %.1f us, decision published %.1f us | a leader warp needed it at %.1f us\n",
                            v, c, (q[2] - t0) * 1e-3, (q[3] - t0) * 1e-3, (q[0] - t0) * 1e-3, (q[1] - t0) * 1e-3, (need[(size_t)c * 128 + v] - t0) * 1e-3);
                }
        }
        if (getenv("CGG_PROFILE_WARPS")) {
            std::vector<unsigned long long> pw(2 * (size_t)d.G * NWARPS);
            CK(cudaMemcpy(pw.data(), h->prof_dev + 32 + 4096 + 32 * 128 * 4, 8 * pw.size(), cudaMemcpyDeviceToHost));
            for (int b = 0; b < d.G; b += 1) {
                fprintf(stderr, "[cgg warps] cta %3d wait(k cycles):", b);
                for (int w = 0; w < NWARPS; ++w) fprintf(stderr, " %5.0f", pw[2 * ((size_t)b * NWARPS + w)] * 1e-3);
                fprintf(stderr, " | finish of (chain 0, pass 60), us after the first:");
                unsigned long long tmin = ~0ULL;
                for (size_t i = 1; i < pw.size(); i += 2) if (pw[i] && pw[i] < tmin) tmin = pw[i];
                for (int w = 0; w < NWARPS; ++w) fprintf(stderr, " %4.1f", (pw[2 * ((size_t)b * NWARPS + w) + 1] - tmin) * 1e-3);
                fprintf(stderr, "\n");
            }
        }
        if (getenv("CGG_PROFILE_CTAS")) {
            std::vector<unsigned long long> pc(4 * (size_t)d.G);
            CK(cudaMemcpy(pc.data(), h->prof_dev + 32, 8 * pc.size(), cudaMemcpyDeviceToHost));
            for (int b = 0; b < d.G; ++b)
                fprintf(stderr, "[cgg cta] %3d sm %3llu wait %.3g tiles %.3g rows %.3g\n", b, pc[4 * b + 3], (double)pc[4 * b] / NWARPS, (double)pc[4 * b + 1] / NWARPS, (double)pc[4 * b + 2] / NWARPS);
        }
    }

    CK(cudaMemcpy(&hdr, d.hdr, sizeof hdr, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cs.data(), d.cs, sizeof(ChainState) * C, cudaMemcpyDeviceToHost));
    if (samples_out) CK(cudaMemcpy(samples_out, d.samples, sizeof(double) * ns, cudaMemcpyDeviceToHost));
    cgg_stats st;
    memset(&st, 0, sizeof st);
    int bad = CGG_OK, bad_chain = -1;
    for (int c = 0; c < C; ++c) {
        st.updates += cs[c].updates; st.chain_passes += cs[c].chain_passes; st.commit_passes += cs[c].commit_passes;
        st.cand_evals += cs[c].cand_evals; st.ref_evals += cs[c].ref_evals; st.stepouts += cs[c].stepouts; st.shrinks += cs[c].shrinks;
        st.passes += cs[c].passes; st.coarse_evals += cs[c].coarse_evals; st.coarse_undecided += cs[c].coarse_undecided;
        st.jet_passes += cs[c].jet_passes; st.jet_fallbacks += cs[c].jet_fallbacks; st.jet_retries += cs[c].jet_retries;
        if (u_consumed) u_consumed[c] = cs[c].cursor;
        if (cs[c].status != CGG_OK && bad == CGG_OK) { bad = cs[c].status; bad_chain = c; }
    }
    if (d.jet_light && d.family == CGG_BINOMIAL) std::fill(h->fx_valid.begin(), h->fx_valid.end(), 0);   // carried f(x0) is a surrogate value
    h->last_cs = cs;
    if (mailboxes) {       // stamps are per-chain pass numbers offset by a common base: move the base past every chain's count
        uint64_t mx = 0;
        for (int c = 0; c < C; ++c) mx = std::max<uint64_t>(mx, cs[c].passes);
        h->mbox_stamp += (uint32_t)mx + 2u;
    }
    // a chain that failed on the device stopped in the middle of an update (its beta may be ahead of its eta): it has to
    // be initialised again before it can run (cgg_init_chain / cgg_set_state)
    for (int c = 0; c < C; ++c)
        if (cs[c].status != CGG_OK || hdr.abort) { h->chain_init[c] = 0; h->fx_valid[c] = 0; h->fx_mag[c] = 0; }
    st.launches = launches; st.sweep_ms = ms; st.group_passes = (uint64_t)hdr.group_passes;
    st.algorithmic_bytes = 8.0 * (double)d.n * (3.0 * (double)st.chain_passes + 2.0 * (double)st.commit_passes);
    if (stats) *stats = st;
    if (hdr.abort) return fail(CGG_E_CUDA, "cgg_run: a chain wait timed out (a worker never arrived)");
    if (bad != CGG_OK) {
        const char *why = bad == CGG_E_NAN ? "log-potential is NaN (the reference would stop with 'missing value where TRUE/FALSE needed')"
                        : bad == CGG_E_STREAM ? "replay-uniform stream exhausted"
                        : bad == CGG_E_NOTERM ? "slice loop did not terminate" : "device-side failure";
        return fail(bad, "cgg_run: chain %d: %s", bad_chain, why);
    }
    return CGG_OK;
}
