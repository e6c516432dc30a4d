// ptl_advance_wf.cuh — K1 "wavefront" variant of the fused advance kernel for collision-dominated
// species (electrons, positrons, slow electrons: kappa >> 1 sub-steps per particle per dt).
//
// Why: with one particle per thread and the reference's control flow (mixed_population.jl:61-88,
// collisions.jl:142-199) a warp executes the UNION of what its 32 lanes do — push, 15-way process
// switch, rejection loops with ~27 % acceptance (RBEB) — and ncu measured 6.6 active lanes per
// instruction (profiles/r1_v1_electron_ncu_summary.csv).  Here a persistent CTA keeps the state of
// WF_THREADS particles in SHARED MEMORY and splits the per-particle loop into short work units
//     STEP (push to next event + null-collision selection) | COULOMB (sample + apply) |
//     RBEB (rejection trials) | IONFIN (two-body kinematics + apply + birth) | OTHER (rare processes)
// Every round the CTA counting-sorts its slots by pending unit (warp ballots + a shared-memory
// prefix), so thread k executes the k-th unit of the sorted order and warps are coherent; finished
// slots are refilled from a global row counter (dynamic load balance, no tail inside the block).
// The arithmetic, the draw order and the per-particle Philox stream are exactly those of the
// one-thread-per-particle kernel, so results are identical particle by particle.
#pragma once
#include "ptl_advance.cuh"

namespace ptl {

constexpr int WF_THREADS = 256;
constexpr int WF_WARPS = WF_THREADS / 32;
#ifndef WF_MIN_BLOCKS
#define WF_MIN_BLOCKS 2
#endif

// (class 4 was a separate IONFIN unit in round 1; it is the second half of the asynchronous LOAD now)
enum { WS_LOAD = 0, WS_STEP, WS_COULOMB, WS_RBEB, WS_LOADWAIT, WS_OTHER, WS_IDLE, WS_NCLASS };
static_assert(WS_IDLE == 6 && WF_WARPS == 8, "the lane-parallel scheduler assumes 6 work classes x 8 warps");
constexpr int WF_CUM_STRIDE = 50;        // doubles per energy interval of the shared-memory cumulative-rate table
#ifdef WF_LINEAR_SELECT
#define WF_CUM_PAD (-INFINITY)           // full scan: padded processes never win
#else
#define WF_CUM_PAD INFINITY              // binary search: padded processes sort last
#endif
#ifndef WF_RBEB_TRIALS
#define WF_RBEB_TRIALS 0                 // rejection trials evaluated side by side per RBEB unit (0: the sequential two-trial loop)
#endif
#ifndef WF_RBEB_DYN_MIN
#define WF_RBEB_DYN_MIN 0
#endif
#ifndef WF_QUAD_SELECT
#define WF_QUAD_SELECT 0
#endif
#ifndef WF_STEP_PHILOX_CALL
#define WF_STEP_PHILOX_CALL 1               // 1: the STEP unit calls philox_block like every other unit (62 instructions less hot code: main pass 25.2 -> 24.65 ms);
                                         // 0: the ten rounds inline (round 1: overlapped with the push; measured slower now)
#endif
#ifndef WF_STEP_LOG_CALL
#define WF_STEP_LOG_CALL 0
#endif
#ifndef WF_RBEB_LOOP
#define WF_RBEB_LOOP 5                   // sequential rejection trials per RBEB unit before the slot goes back to the scheduler (see the RBEB unit)
#endif
constexpr uint32_t WF_VALID = 0x100u;   // slot holds a particle that must be written back
constexpr uint32_t WF_DEAD = 0x200u;    // ... and it was deactivated
constexpr uint32_t WF_COAST = 0x400u;   // OTHER unit: no collision, take the repeated below-cut sub-steps in blocks

// double columns of the shared-memory particle pool
// (round 1 also kept the collision energy and the sampled E2 here; they are recomputed / passed in registers now: 16 bytes
// per slot buy a third slot per lane in the warp-private kernel at two CTAs per SM)
enum { WD_X0 = 0, WD_X1, WD_X2, WD_P0, WD_P1, WD_P2, WD_T, WD_S, WD_R, WD_TREM, WD_NCOL };

struct WfPool {
    double* d;              // [WD_NCOL][np]
    int np;                 // slots in the pool (column stride)
    unsigned long long* uid;
    long long* row;
    uint32_t *idx, *cblock, *c2, *c3, *state;   // state: class | flags | (proc index << 16)
    unsigned short* order;
    uint32_t* cnt;          // [class][WF_WARPS], 64 entries
};

// doubles + uid/row + 5 u32 arrays + class-count matrix (64 entries) + order (u16), rounded up to 16 bytes
constexpr size_t WF_POOL_BYTES =
    ((sizeof(double) * WD_NCOL * WF_THREADS + 16 * WF_THREADS + 4 * 5 * WF_THREADS + 4 * 64 + 2 * WF_THREADS) + 15) / 16 * 16;
__host__ __device__ inline size_t wf_pool_bytes() { return WF_POOL_BYTES; }

__device__ __forceinline__ void wf_load_rng(const WfPool& S, int it, Rng& rng) {
    unsigned long long uid = S.uid[it];
    rng.k0 = (uint32_t)uid;
    rng.k1 = (uint32_t)(uid >> 32) ^ DOM_COLLISION;
    rng.idx = S.idx[it]; rng.cblock = S.cblock[it]; rng.c2 = S.c2[it]; rng.c3 = S.c3[it];
}
__device__ __forceinline__ void wf_store_rng(const WfPool& S, int it, const Rng& rng) {
    S.idx[it] = rng.idx; S.cblock[it] = rng.cblock; S.c2[it] = rng.c2; S.c3[it] = rng.c3;
}
__device__ __forceinline__ Vec3 wf_get3(const WfPool& S, int c0, int it) {
    return {S.d[(c0 + 0) * S.np + it], S.d[(c0 + 1) * S.np + it], S.d[(c0 + 2) * S.np + it]};
}
__device__ __forceinline__ void wf_put3(const WfPool& S, int c0, int it, Vec3 v) {
    S.d[(c0 + 0) * S.np + it] = v.x; S.d[(c0 + 1) * S.np + it] = v.y; S.d[(c0 + 2) * S.np + it] = v.z;
}
#define WFD(col, it) S.d[(col) * S.np + (it)]

// A particle that the field decelerated below energy_cut keeps its old r != 0: do_one_collision! returns before it
// draws anything (collisions.jl:148-151), so the sub-step loop of mixed_population.jl:66-87 repeats the same
// dt = s/r push until tfinal or until the energy is back above the cut -- up to 1e5..1e6 sequential sub-steps when s
// is small, which used to be the tail of every launch.  Under the uniform-E fast path the kinetic energy is a convex
// function of time, so "below the cut after j sub-steps" implies below the cut at every sub-step in between: take the
// j sub-steps as ONE RK2 step of j*dt (p is linear in t and exact up to rounding, x differs by the RK2 truncation error
// of the longer step, < 1e-10 relative here), with t and trem accumulated sub-step by sub-step so that they stay
// bit-identical, and halve j when the energy would cross the cut so that the first test at or above the cut is
// still taken by the regular path on the reference's own time grid.
// Works on the slot in the shared-memory pool (it runs as a flagged OTHER unit, so the hot STEP unit pays one compare)
// and takes at most ~256 sub-steps per call so that the unit stays as short as the others in its round; the slot keeps
// the WF_COAST flag until no further repeated sub-step fits or the next one would reach the cut.
template <int SP>
static __device__ __noinline__ unsigned long long wf_coast_below_cut(const AdvanceParams& P, const WfPool& S, int it, double cut) {
    Vec3 x = wf_get3(S, WD_X0, it), p = wf_get3(S, WD_P0, it);
    double t = WFD(WD_T, it), trem = WFD(WD_TREM, it);
    const double tnext = WFD(WD_S, it) * frcp(WFD(WD_R, it));       // same expression as the STEP unit (:67)
    bool more = false;
    const unsigned long long nsub = coast_below_cut<SP>(P, x, p, t, trem, tnext, cut, 256, more);
    if (nsub) {
        wf_put3(S, WD_X0, it, x); wf_put3(S, WD_P0, it, p);
        WFD(WD_T, it) = t; WFD(WD_TREM, it) = trem;
    }
    S.state[it] = more ? (WS_OTHER | WF_VALID | WF_COAST) : (WS_STEP | WF_VALID);
    return nsub;
}

// after a real collision changed p: apply!'s setr! (collisions.jl:93,99) and back to STEP (wf_setr_cheb3: ptl_advance.cuh)
template <int SP>
__device__ __forceinline__ void wf_after_collision(const AdvanceParams& P, const SmemTable& T, const WfPool& S, int it, Vec3 p, double s,
                                                   const bool cheb3 = false, const double cut = 0.0) {
    wf_put3(S, WD_P0, it, p);
    WFD(WD_S, it) = s;
    WFD(WD_R, it) = cheb3 ? wf_setr_cheb3<SP>(P, P.tab[SP], T.ratebound, cut, p) : setr<SP>(P, T, p);
    S.state[it] = WS_STEP | WF_VALID;
}

// process selection of do_one_collision! (collisions.jl:154-196).  The reference scans the processes
// sequentially (xi -= nu_j until nu_j > xi); in exact arithmetic that picks the first j whose running sum
// cum_j exceeds xi0 = u*r.  Lanes stop at different j (and null events scan all of them), so the scan is the
// most divergent part of a STEP unit.  Here every lane evaluates all cum_j rows (uniform trip count, full
// lanes) and takes the first j with cum_j > xi0; whenever some |cum_j - xi0| is within a guard band that is
// orders of magnitude wider than the rounding differences between the two formulations, the lane falls back
// to the reference's sequential scan on the per-process table, so the selected process is always identical.
template <int TK, bool FAST>
__device__ __forceinline__ int wf_select(const AdvanceParams& P, const TableView& T, const double* __restrict__ cum, const Pre& pre,
                                         double xi0, double r, bool& rate_bound_violated);

// cold paths of the selection (tables that do not fit the fast layout; exact sequential scan inside the guard band)
template <int TK>
static __device__ __noinline__ int wf_select_generic(const AdvanceParams& P, const TableView& T, const double* __restrict__ cum, Pre pre,
                                              double xi0, double r, int* rbv) {
    bool b;
    int j = wf_select<TK, false>(P, T, cum, pre, xi0, r, b);
    *rbv = b ? 1 : 0;
    return j;
}
template <int TK>
static __device__ __noinline__ int wf_select_sequential(const TableView& T, Pre pre, double xi0, int* rbv) {
    const int np = T.nprocs;
    double xi = xi0;
    int jsel = -1;
    for (int j = 0; j < np; j++) {
        double nu = (TK == 0) ? chebsum(T.rate + (size_t)T.order * (j + (size_t)np * pre.i), pre, T.order)
                              : linear_rate(T.rate, np, j, pre);
        if (nu > xi) { jsel = j; break; }
        xi -= nu;
    }
    *rbv = (jsel < 0 && !(xi >= 0)) ? 1 : 0;
    return jsel;
}

template <int TK, bool FAST>
__device__ __forceinline__ int wf_select(const AdvanceParams& P, const TableView& T, const double* __restrict__ cum, const Pre& pre,
                                         double xi0, double r, bool& rate_bound_violated) {
    const int np = T.nprocs;
    int jsel = -1;
    bool nearb = false;
    const double guard = 1e-9 * r;
#ifndef WF_LINEAR_SELECT
    if (TK == 0 && FAST) {
        // Binary search over the running sums.  On an interval where every fitted rate is non-negative (T.mono_mask, checked
        // on the host at table upload) cum_0 <= cum_1 <= ... up to rounding, so "first j with cum_j > xi0" is a lower-bound
        // search: four probes instead of 16 evaluations, then the two entries that bracket xi0 are checked against the guard
        // band -- any disagreement with the reference's sequential subtraction needs one of those two within rounding of xi0.
        // Layout [interval][m][16], padded with +inf (sorts last); np <= 16.
        const double* c = cum + WF_CUM_STRIDE * pre.i;
        const double ca = pre.a, cb = pre.b;
#define WF_CUMJ(j) fma(c[32 + (j)], cb, fma(c[16 + (j)], ca, c[(j)]))
#if WF_QUAD_SELECT
        // Two levels of THREE independent probes (entries 3, 7, 11, then the first three of the quarter that holds xi0)
        // instead of four dependent ones + the two guard-band entries: the same six evaluations, the same bits, two
        // shared-memory round trips on the warp's critical path instead of six.  Measured 15 % SLOWER (28.5 against 24.8 ms,
        // main pass of 4e6 electrons; the full 16-entry scan WF_LINEAR_SELECT: 28.5 as well) — and not for lack of registers:
        // with 12 warps x 168 registers it is 29.2 against 25.3 ms.  The select logic on six live values costs more than the
        // four saved round trips, which other warps were hiding anyway.
        const double e3 = WF_CUMJ(3), e7 = WF_CUMJ(7), e11 = WF_CUMJ(11);
        const int q4 = (!(e3 > xi0) ? 4 : 0) + (!(e7 > xi0) ? 4 : 0) + (!(e11 > xi0) ? 4 : 0);      // monotone: 0, 4, 8 or 12
        const double f0 = WF_CUMJ(q4), f1 = WF_CUMJ(q4 + 1), f2 = WF_CUMJ(q4 + 2);
        const int inq = (!(f0 > xi0) ? 1 : 0) + (!(f1 > xi0) ? 1 : 0) + (!(f2 > xi0) ? 1 : 0);     // entries of the quarter that are <= xi0
        int j = q4 + inq;                                      // entries 0..14 examined: j = how many are <= xi0
        // the entry at j (first one above xi0) and the one below it are among the evaluated ones, except entry 15
        const double top = q4 == 0 ? e3 : (q4 == 4 ? e7 : e11);                 // entry q4 + 3 (only used when q4 < 12)
        double dj = (inq == 0 ? f0 : (inq == 1 ? f1 : (inq == 2 ? f2 : (q4 < 12 ? top : WF_CUMJ(15))))) - xi0;
        const bool over = np > 15 && j == 15 && !(dj > 0);                      // all sixteen entries are <= xi0: null collision
        const double below = q4 == 4 ? e3 : (q4 == 8 ? e7 : e11);               // entry q4 - 1 (only used when q4 > 0)
        const double dp = over ? dj : (inq == 0 ? (q4 > 0 ? below : f0) : (inq == 1 ? f0 : (inq == 2 ? f1 : f2))) - xi0;
        if (over) { j = 16; dj = INFINITY; }
        nearb = (fabs(dj) < guard) | (fabs(dp) < guard) | !((T.mono_mask >> pre.i) & 1ULL);
        jsel = j < np ? j : -1;
#else
        int j = !(WF_CUMJ(7) > xi0) ? 8 : 0;
        j += !(WF_CUMJ(j + 3) > xi0) ? 4 : 0;
        j += !(WF_CUMJ(j + 1) > xi0) ? 2 : 0;
        j += !(WF_CUMJ(j) > xi0) ? 1 : 0;                       // entries 0..14 examined: j = how many are <= xi0
        double dj = WF_CUMJ(j) - xi0;                            // j == 15: the 16th entry (a process or the +inf pad)
        if (np > 15 && j == 15 && !(dj > 0)) j = 16;
        const double dp = WF_CUMJ(j > 0 ? j - 1 : 0) - xi0;
        if (j == 16) dj = INFINITY;
        nearb = (fabs(dj) < guard) | (fabs(dp) < guard) | !((T.mono_mask >> pre.i) & 1ULL);
        jsel = j < np ? j : -1;
#endif
#undef WF_CUMJ
    } else if (TK == 0) {
#else
    if (TK == 0 && FAST) {
        // shared-memory layout [interval][m][16] (order 3, <= 16 processes, padded with -inf): two processes per LDS.128.
        // Interval stride WF_CUM_STRIDE = 50 doubles (400 B = 4 banks mod 32): lanes in different energy intervals hit
        // different banks (a 384-B stride put every interval on the same banks: 14e9 conflict wavefronts in ncu).
        const double2* c0 = reinterpret_cast<const double2*>(cum + WF_CUM_STRIDE * pre.i);
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++) {
            if (2 * j2 >= np) break;
            double2 a0 = c0[j2], a1 = c0[8 + j2], a2 = c0[16 + j2];
            double d0 = fma(a2.x, pre.b, fma(a1.x, pre.a, a0.x)) - xi0;
            double d1 = fma(a2.y, pre.b, fma(a1.y, pre.a, a0.y)) - xi0;
            nearb |= (fabs(d0) < guard) | (fabs(d1) < guard);
            jsel = (jsel < 0 && d0 > 0) ? 2 * j2 : jsel;
            jsel = (jsel < 0 && d1 > 0) ? 2 * j2 + 1 : jsel;
        }
    } else if (TK == 0) {
#endif
        const double* c = cum + T.order * np * pre.i;
        for (int j = 0; j < np; j++) {
            const double* a = c + T.order * j;
            double cj = a[0];
            if (T.order > 1) cj = fma(a[1], pre.a, cj);
            if (T.order > 2) cj = fma(a[2], pre.b, cj);
            if (T.order > 3) {
                double tm2 = pre.a, tm1 = pre.b;
                for (int m = 3; m < T.order; m++) {
                    double tm = 2 * pre.a * tm1 - tm2;
                    cj = fma(a[m], tm, cj);
                    tm2 = tm1; tm1 = tm;
                }
            }
            double d = cj - xi0;
            nearb |= fabs(d) < guard;
            if (jsel < 0 && d > 0) jsel = j;
        }
#ifdef WF_LINEAR_TABLE_SCAN
    } else if (false) {
#else
    } else if (T.mono_mask != 0) {
#endif
        // Linear (LXCat) tables: every tabulated rate is >= 0 (checked on the host), so the interpolated running sums are
        // non-decreasing in j and the first j with cum_j > xi0 is a lower-bound search: ~log2(np) probes of two loads each
        // instead of np (a real N2/O2 set has 50-80 channels and most sub-steps end in the explicit null row).
        const double2* __restrict__ cp = T.cum2 + (size_t)np * pre.i;     // {row i, row i + 1} pairs: one LDG.128 per probe
        const double w1 = 1 - pre.a;
        double2 cv_;
#define WF_CUML(j) (cv_ = __ldg(cp + (j)), pre.a * cv_.x + w1 * cv_.y)
        // (An 8-ary search -- eight independent probes per level, two dependent round trips instead of six for 64 channels --
        // was measured 1.8x SLOWER: 12.7 vs 6.9 ms.  The probes are uncoalesced 8-byte loads, 32 sectors per instruction; the
        // path is bound by that sector traffic, so fewer loads beat shorter chains.)
        int lo = 0, n = np;
        while (n > 0) {
            const int half = n >> 1, mid = lo + half;
            if (!(WF_CUML(mid) > xi0)) { lo = mid + 1; n -= half + 1; } else n = half;
        }
        if (lo < np) { nearb |= fabs(WF_CUML(lo) - xi0) < guard; jsel = lo; }
        if (lo > 0) nearb |= fabs(WF_CUML(lo - 1) - xi0) < guard;
#undef WF_CUML
    } else {
        for (int j = 0; j < np; j++) {
            double c0 = __ldg(cum + j + (size_t)np * pre.i), c1 = __ldg(cum + j + (size_t)np * (pre.i + 1));
            double d = (pre.a * c0 + (1 - pre.a) * c1) - xi0;
            nearb |= fabs(d) < guard;
            if (jsel < 0 && d > 0) jsel = j;
        }
    }
    rate_bound_violated = false;
    if (nearb) {   // exact sequential scan of the reference on the per-process rates (global memory; ~never taken)
        int rbv = 0;
        jsel = wf_select_sequential<TK>(T, pre, xi0, &rbv);
        rate_bound_violated = rbv != 0;
    }
    return jsel;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}


// The OTHER unit (rare processes: the whole collide() + apply!; also the block-coasting of below-cut particles) as a
// function of its own.  Inlined it is ~2000 instructions of cold code in the middle of the kernel body, between the STEP
// and the RBEB units.  Out of line (WQ_OTHER_OUTLINE, warp-private kernel) it takes scalars only and rebuilds the pool and
// table views from the shared-memory base, so that no pointer has to stay live in the caller for its sake.
template <int SP, int TK>
__device__ __forceinline__ unsigned long long wf_other_unit(const AdvanceParams& P, const PopView& Q, const WfPool& S, const SmemTable& TS,
                                                            const bool fastsel, const RngCtx rc, const double cut, const int it, const uint32_t sw) {
    unsigned long long nsub = 0;
    do {
        if (sw & WF_COAST) {
            nsub += wf_coast_below_cut<SP>(P, S, it, cut);
            break;
        }
        Rng rng;
        wf_load_rng(S, it, rng);
        Vec3 p = wf_get3(S, WD_P0, it), x = wf_get3(S, WD_X0, it);
        double eng = kinenergy<SP>(p), t = WFD(WD_T, it);       // same p, same function as the STEP unit's test: same bits
        Outcome o;
        collide<SP>(rng, rc, P, TS.procs[sw >> 16], p, eng, o);
        wf_store_rng(S, it, rng);
        long long i = S.row[it];
        uint64_t cu[2];
        switch (o.kind) {
        case OUT_NULL:
            WFD(WD_R, it) = setr<SP>(P, TS, p);
            WFD(WD_S, it) = -nlog(rng.u(rc.step, rc.seed_lo, rc.seed_hi));
            wf_store_rng(S, it, rng);
            S.state[it] = WS_STEP | WF_VALID;
            break;
        case OUT_STATE_CHANGE:
            wf_after_collision<SP>(P, TS, S, it, o.p1, o.s1, fastsel, cut);
            break;
        case OUT_NEW_PARTICLE:
            wf_after_collision<SP>(P, TS, S, it, o.p1, o.s1, fastsel, cut);
            child_uids(S.uid[it], rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
            add_particle(P, o.sp2, x, o.p2, Q.col[COL_W][i], t, o.s2, cu[0]);
            break;
        case OUT_REMOVE:
            S.state[it] = WS_LOAD | WF_VALID | WF_DEAD;
            break;
        case OUT_REPLACE:
            S.state[it] = WS_LOAD | WF_VALID | WF_DEAD;
            child_uids(S.uid[it], rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
            add_particle(P, o.sp2, x, o.p2, Q.col[COL_W][i], t, o.s2, cu[0]);
            break;
        case OUT_REPLACE_PAIR:
            S.state[it] = WS_LOAD | WF_VALID | WF_DEAD;
            child_uids(S.uid[it], rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
            add_particle(P, o.sp2, x, o.p2, Q.col[COL_W][i], t, o.s2, cu[0]);
            add_particle(P, o.sp3, x, o.p3, Q.col[COL_W][i], t, o.s3, cu[1]);
            break;
        }
    } while (0);
    return nsub;
}

// The LOAD unit (write back the finished particle of a slot, fetch the next row) as a function of its own: 1 % of the
// rounds, ~250 instructions that otherwise sit between the STEP and the RBEB units in the kernel body.
template <int SP, int TK, bool FIRST, bool ALOAD>
__device__ __forceinline__ void wf_load_unit(const AdvanceParams& P, const TableView& T, const PopView& Q, const WfPool& S, const SmemTable& TS,
                                             const bool fastsel, const double cut, const int it, const uint32_t sw, const unsigned ldmask,
                                             const int lane, const unsigned ltmask, unsigned long long* row_counter, const long long i0,
                                             const long long i1, const long long* __restrict__ rows) {
    do {
        if (sw & WF_VALID) {
            long long i = S.row[it];
            Q.col[COL_X0][i] = WFD(WD_X0, it); Q.col[COL_X1][i] = WFD(WD_X1, it); Q.col[COL_X2][i] = WFD(WD_X2, it);
            Q.col[COL_P0][i] = WFD(WD_P0, it); Q.col[COL_P1][i] = WFD(WD_P1, it); Q.col[COL_P2][i] = WFD(WD_P2, it);
            Q.col[COL_T][i] = WFD(WD_T, it); Q.col[COL_S][i] = WFD(WD_S, it); Q.col[COL_R][i] = WFD(WD_R, it);
            if (sw & WF_DEAD) Q.active[i] = 0;
        }
        int leader = __ffs(ldmask) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(row_counter, (unsigned long long)__popc(ldmask));
        base = __shfl_sync(ldmask, base, leader);
        long long i = i0 + (long long)base + __popc(ldmask & ltmask);
        if (i >= i1) { S.state[it] = WS_IDLE; break; }
        if (rows != nullptr) i = rows[i];
        if (ALOAD) {
            cp_async8(&WFD(WD_X0, it), Q.col[COL_X0] + i); cp_async8(&WFD(WD_X1, it), Q.col[COL_X1] + i); cp_async8(&WFD(WD_X2, it), Q.col[COL_X2] + i);
            cp_async8(&WFD(WD_P0, it), Q.col[COL_P0] + i); cp_async8(&WFD(WD_P1, it), Q.col[COL_P1] + i); cp_async8(&WFD(WD_P2, it), Q.col[COL_P2] + i);
            cp_async8(&WFD(WD_T, it), Q.col[COL_T] + i); cp_async8(&WFD(WD_S, it), Q.col[COL_S] + i);
            if (!FIRST) cp_async8(&WFD(WD_R, it), Q.col[COL_R] + i);
            cp_async8(&S.uid[it], Q.uid + i);
            cp_async4(&S.c2[it], Q.active + (i & ~3LL));              // the aligned word that holds the row's active flag
            asm volatile("cp.async.commit_group;" ::: "memory");
            S.row[it] = i;
            S.state[it] = WS_LOADWAIT;
            break;
        }
        if (!Q.active[i]) { S.state[it] = WS_LOAD; break; }    // l.active || continue  (mixed_population.jl:63)
        Vec3 x = {Q.col[COL_X0][i], Q.col[COL_X1][i], Q.col[COL_X2][i]};
        Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
        double t = Q.col[COL_T][i];
        wf_put3(S, WD_X0, it, x); wf_put3(S, WD_P0, it, p);
        WFD(WD_T, it) = t; WFD(WD_S, it) = Q.col[COL_S][i];
        WFD(WD_R, it) = FIRST ? (fastsel ? wf_setr_cheb3<SP>(P, T, TS.ratebound, cut, p) : setr<SP>(P, TS, p))
                              : Q.col[COL_R][i];                           // advance_init!  mixed_population.jl:97-110
        WFD(WD_TREM, it) = P.tfinal - t;                                   // :65
        S.uid[it] = Q.uid[i]; S.row[it] = i;
        S.idx[it] = 0; S.cblock[it] = 0xFFFFFFFFu; S.c2[it] = 0; S.c3[it] = 0;
        S.state[it] = WS_STEP | WF_VALID;
            } while (0);
}

// One work unit of slot `it` (state word `sw`).  Shared by the barrier-synchronous kernel (k_advance_wf) and the
// queue-driven kernel (k_advance_aq).  `ldmask` = lanes of this warp that execute a LOAD unit right now (they share one
// atomic on the global row counter).
// ALOAD (warp-private kernel only): the LOAD unit does not wait for the row — it issues cp.async copies of the twelve
// column entries straight into the slot and parks the slot in class LOADWAIT; the warp goes on with other classes while
// HBM answers (the synchronous LOAD was the one long-scoreboard stall of the kernel: 0.55-0.7 warps per issue), and the
// LOADWAIT unit, run after a warp-wide cp.async.wait_all, finishes advance_init! on the arrived row.
template <int SP, int TK, bool FIRST, bool CB, bool ALOAD = false>
__device__ __forceinline__ void wf_execute_unit(const AdvanceParams& P, const TableView& T, const PopView& Q, const WfPool& S,
                                                const SmemTable& TS, const double* tcum, const bool fastsel, const RngCtx rc,
                                                const double cut, const int it, const uint32_t sw, const unsigned ldmask,
                                                const int lane, const unsigned ltmask, unsigned long long* row_counter,
                                                const long long i0, const long long i1, unsigned long long& nsub,
                                                const long long* __restrict__ rows = nullptr, const int cls_uniform = -1,
                                                unsigned long long (*other_out)(const AdvanceParams*, int, uint32_t) = nullptr,
                                                void (*load_out)(const AdvanceParams*, int, uint32_t, unsigned, unsigned long long*, long long, long long, const long long*) = nullptr) {
    // cls_uniform >= 0: the caller guarantees that every executing lane of the warp holds a slot of this class (the
    // warp-private kernel), so the dispatch is a warp-uniform branch instead of a divergent switch on the state word
    const int cls = cls_uniform >= 0 ? cls_uniform : (int)(sw & 0xffu);
    switch (cls) {
    // ------------------------------------------------------------------------------------------
    case WS_LOAD: {   // write back the finished particle of this slot (if any), fetch the next row
        if (load_out != nullptr) load_out(&P, it, sw, ldmask, row_counter, i0, i1, rows);
        else wf_load_unit<SP, TK, FIRST, ALOAD>(P, T, Q, S, TS, fastsel, cut, it, sw, ldmask, lane, ltmask, row_counter, i0, i1, rows);
        break;
    }
    // ------------------------------------------------------------------------------------------
    case WS_LOADWAIT: {   // the row has arrived in the slot (the caller waited for the warp's cp.async groups)
        const long long i = S.row[it];
        if (((S.c2[it] >> (8 * (int)(i & 3))) & 0xffu) == 0u) { S.state[it] = WS_LOAD; break; }    // l.active || continue  (:63)
        if (FIRST) {
            const Vec3 p = wf_get3(S, WD_P0, it);
            WFD(WD_R, it) = fastsel ? wf_setr_cheb3<SP>(P, T, TS.ratebound, cut, p) : setr<SP>(P, TS, p);   // advance_init!  :97-110
        }
        WFD(WD_TREM, it) = P.tfinal - WFD(WD_T, it);                       // :65
        S.idx[it] = 0; S.cblock[it] = 0xFFFFFFFFu; S.c2[it] = 0; S.c3[it] = 0;
        S.state[it] = WS_STEP | WF_VALID;
        break;
    }
    // ------------------------------------------------------------------------------------------
    case WS_STEP: {   // one iteration of mixed_population.jl:66-87 up to the process selection
        double trem = WFD(WD_TREM, it);
        if (!(trem > DBL_EPS)) { S.state[it] = WS_LOAD | WF_VALID; break; }       // :66
        double s = WFD(WD_S, it), r = WFD(WD_R, it), t = WFD(WD_T, it);
        Vec3 x = wf_get3(S, WD_X0, it), p = wf_get3(S, WD_P0, it);
        double tnext = r == 0.0 ? INFINITY : s * frcp(r);   // :67  (r == 0 -> Inf -> free flight)
        bool collides = trem > tnext;                   // :68
        double dt = collides ? tnext : trem;
        if (!collides) s -= dt * r;                     // :74
        // Speculative draws (no in-loop callbacks): the Philox block of this sub-step's collision test depends only on the
        // slot's RNG cursor and the null outcome's s = -log(u2) only on that block, so both are computed HERE, inline, where
        // their ~220 cycles of dependent integer / polynomial latency overlap the push, the energy and the table search
        // instead of following them through two out-of-line calls.  The warp executed both anyway whenever one lane had a
        // null event.  The cursor is committed only if the test really draws (same stream, same values as rng.u()).
        uint32_t sp_idx = 0, sp_o2 = 0, sp_o3 = 0;
        double sp_u1 = 0.0, sp_snull = 0.0;
        if (!CB) {
            const unsigned long long uid = S.uid[it];
            const uint32_t ic = S.idx[it];
            sp_idx = ic + (ic & 1u);                    // every collision test starts on an even draw index
            uint32_t o4[4];
#if WF_STEP_PHILOX_CALL      // the out-of-line copy every other unit uses: 62 instructions less hot code, one call more on the chain
            { const uint4 ob = philox_block(sp_idx >> 1, rc.step, rc.seed_lo, rc.seed_hi, (uint32_t)uid, (uint32_t)(uid >> 32) ^ DOM_COLLISION);
              o4[0] = ob.x; o4[1] = ob.y; o4[2] = ob.z; o4[3] = ob.w; }
#else
            philox4x32_10(sp_idx >> 1, rc.step, rc.seed_lo, rc.seed_hi, (uint32_t)uid, (uint32_t)(uid >> 32) ^ DOM_COLLISION, o4);
#endif
            sp_u1 = bits_to_u01(o4[0], o4[1]);
#if WF_STEP_LOG_CALL
            sp_snull = -nlog(bits_to_u01(o4[2], o4[3]));
#else
            sp_snull = -flog_u01(bits_to_u01(o4[2], o4[3]));
#endif
            sp_o2 = o4[2]; sp_o3 = o4[3];
        }
        Vec3 xo = x, po = p;
        double to = t;
        push<SP>(P, x, p, t, dt);                       // :77
        trem -= dt;                                     // :86
        nsub++;
        // a free flight ends the step (trem - dt == 0); after a collision the loop test is re-evaluated
        uint32_t next = collides ? (WS_STEP | WF_VALID) : (WS_LOAD | WF_VALID);
        bool act = true;
        Rng rng;
        bool rng_loaded = false;
        if (CB) {                                       // onadvance(WallCallback)  callback.jl:167-184
            for (int k = 0; k < P.cb.nwalls; k++) {
                const ptl_wall_desc& wd = P.cb.wall[k];
                if (wd.species != SP) continue;
                double xoc = wd.coord == 0 ? xo.x : (wd.coord == 1 ? xo.y : xo.z);
                double xnc = wd.coord == 0 ? x.x : (wd.coord == 1 ? x.y : x.z);
                if (xoc < wd.v && wd.v < xnc) {
                    double f = (wd.v - xoc) / (xnc - xoc);
                    if (!rng_loaded) { wf_load_rng(S, it, rng); rng_loaded = true; }
                    rng.skip();   // lincomb's 4-arg constructor draws (and discards) an s  (electron.jl:127-132)
                    const WallBuf& W = P.wall[k];
                    double wgt = Q.col[COL_W][S.row[it]];
                    unsigned long long slot = atomicAdd(W.n, 1ULL);
                    if ((long long)slot < W.capacity) {
                        W.col[0][slot] = x.x * f + xo.x * (1 - f); W.col[1][slot] = x.y * f + xo.y * (1 - f); W.col[2][slot] = x.z * f + xo.z * (1 - f);
                        W.col[3][slot] = p.x * f + po.x * (1 - f); W.col[4][slot] = p.y * f + po.y * (1 - f); W.col[5][slot] = p.z * f + po.z * (1 - f);
                        W.col[6][slot] = wgt * f + wgt * (1 - f);
                        W.col[7][slot] = t * f + to * (1 - f);
                    } else {
                        atomicOr(P.flags, PTL_ERR_CAPACITY_OVERFLOW);
                    }
                    if (wd.drop) act = false;
                }
            }
        }
        if (!act) next = WS_LOAD | WF_VALID | WF_DEAD;
        if (collides && act) {                          // :83  do_one_collision!  collisions.jl:142-199
            double eng;
            if (r != 0.0 && (eng = kinenergy<SP>(p)) >= cut) {                     // :148-151
                Pre pre = (TK == 0) ? precheb(eng, T.k, T.xmax, T.rxmax) : indweight(T, eng);   // :153
                if (pre.oob) atomicOr(P.flags, PTL_ERR_ENERGY_OUT_OF_TABLE);
                double xi;
                if (CB) {
                    if (!rng_loaded) { wf_load_rng(S, it, rng); rng_loaded = true; }
                    rng.idx += rng.idx & 1u;   // every collision test starts on an even draw index (Philox block boundary)
                    xi = rng.u(rc.step, rc.seed_lo, rc.seed_hi) * r;                // :154
                } else {
                    xi = sp_u1 * r;
                }
                const int np = T.nprocs;
                bool rbv;
                int jsel;                                                                   // :166-180
                        if (fastsel) {
                            jsel = wf_select<TK, true>(P, T, tcum, pre, xi, r, rbv);
                        } else {
                            int rb2 = 0;
                            jsel = wf_select_generic<TK>(P, T, tcum, pre, xi, r, &rb2);
                            rbv = rb2 != 0;
                        }
                if (rbv) atomicOr(P.flags, PTL_ERR_RATE_BOUND_VIOLATED);            // :186
                if (CB && P.cb.count_collisions) atomicAdd(T.counts + (jsel >= 0 ? jsel : np), 1ULL);
                int kind = jsel >= 0 ? TS.procs[jsel].kind : PTL_PROC_NULL;
                if (kind == PTL_PROC_NULL) {            // NullOutcome: setr! then s = nextcoll()  (:83-88, :182-196)
                    // setr! recomputes kinenergy and presample from the same p: reuse them
                    r = (TK == 0) ? chebsum(TS.ratebound + T.order * pre.i, pre, T.order) : (T.rbvec != nullptr ? linear_bound(T, pre) : T.maxrate);
                    if (CB) {
                        s = -nlog(rng.u(rc.step, rc.seed_lo, rc.seed_hi));
                    } else {
                        s = sp_snull;
                        S.idx[it] = sp_idx + 2;         // both halves of the block consumed: nothing to cache
                    }
                } else {
                    if (!CB) {                          // xi consumed; the second half of the block stays cached for the sampler
                        S.idx[it] = sp_idx + 1; S.cblock[it] = sp_idx >> 1; S.c2[it] = sp_o2; S.c3[it] = sp_o3;
                    }
                    uint32_t c = kind == PTL_PROC_COULOMB ? WS_COULOMB : (kind == PTL_PROC_RBEB ? WS_RBEB : WS_OTHER);
                    next = c | WF_VALID | ((uint32_t)jsel << 16);
                }
            } else if (!CB && r != 0.0 && P.fast_force && trem > tnext) {
                next = WS_OTHER | WF_VALID | WF_COAST;  // below the cut with r != 0: see wf_coast_below_cut
            }
        }
        if (rng_loaded) wf_store_rng(S, it, rng);
        wf_put3(S, WD_X0, it, x); wf_put3(S, WD_P0, it, p);
        WFD(WD_T, it) = t; WFD(WD_S, it) = s; WFD(WD_R, it) = r; WFD(WD_TREM, it) = trem;
        S.state[it] = next;
        break;
    }
    // ------------------------------------------------------------------------------------------
    case WS_COULOMB: {   // collide(::RelativisticCoulomb) + apply!(StateChange)
        // (Computing this unit's two Philox blocks, the azimuth's sincospi and -log(s) inline up front, so that they overlap
        // the sampler, was measured 3.5 % SLOWER: 34.0 vs 32.9 ms -- the extra inline code costs more than the overlap gains.)
        Rng rng;
        wf_load_rng(S, it, rng);
        Vec3 p = wf_get3(S, WD_P0, it);
        Outcome o;
        collide_coulomb<SP>(rng, rc, TS.procs[sw >> 16], p, o);
        wf_store_rng(S, it, rng);
        wf_after_collision<SP>(P, TS, S, it, o.p1, o.s1, fastsel, cut);
        break;
    }
    // ------------------------------------------------------------------------------------------
    case WS_RBEB: {   // rejection trials of rbeb.jl:170-196, WF_RBEB_TRIALS per unit
#if WF_RBEB_TRIALS > 0
        // Trial q of the reference's loop consumes draws (2q, 2q+1) after the cursor whatever the earlier trials decided,
        // so the trials of one unit are independent: their Philox blocks and acceptance tests are evaluated side by side
        // (instruction-level parallelism instead of call -> test -> call -> test), the FIRST accepted one wins and the
        // cursor advances only past the draws the sequential loop would have consumed.  Same stream, same values.
        constexpr int NT = WF_RBEB_TRIALS;
        const unsigned long long uid = S.uid[it];
        const uint32_t k0 = (uint32_t)uid, k1 = (uint32_t)(uid >> 32) ^ DOM_COLLISION;
        const uint32_t idx0 = S.idx[it], b0 = idx0 >> 1;
        const bool po = (idx0 & 1u) != 0;
        uint32_t bw[NT + 1][4];
        if (po && S.cblock[it] == b0) { bw[0][0] = 0; bw[0][1] = 0; bw[0][2] = S.c2[it]; bw[0][3] = S.c3[it]; }
        else philox4x32_10(b0, rc.step, rc.seed_lo, rc.seed_hi, k0, k1, bw[0]);
#pragma unroll
        for (int m = 1; m <= NT; m++) philox4x32_10(b0 + m, rc.step, rc.seed_lo, rc.seed_hi, k0, k1, bw[m]);
        double eng = kinenergy<SP>(wf_get3(S, WD_P0, it));      // same p, same function as the STEP unit's test: same bits
        double B = TS.procs[sw >> 16].par[0];
        RbebConsts k = rbeb_consts<true>(eng, B);
        double w = 0.0;
        int qacc = -1;
#pragma unroll
        for (int q = NT - 1; q >= 0; q--) {           // descending: the lowest accepted q is written last
            const int j0 = 2 * q, j1 = 2 * q + 1;       // draw j sits in block (po + j) >> 1, half (po + j) & 1
            const double u = bits_to_u01(po ? bw[(1 + j0) >> 1][2 * ((1 + j0) & 1)] : bw[j0 >> 1][2 * (j0 & 1)],
                                         po ? bw[(1 + j0) >> 1][2 * ((1 + j0) & 1) + 1] : bw[j0 >> 1][2 * (j0 & 1) + 1]);
            const double u2 = bits_to_u01(po ? bw[(1 + j1) >> 1][2 * ((1 + j1) & 1)] : bw[j1 >> 1][2 * (j1 & 1)],
                                          po ? bw[(1 + j1) >> 1][2 * ((1 + j1) & 1) + 1] : bw[j1 >> 1][2 * (j1 & 1) + 1]);
            double wq;
            if (rbeb_trial(k, u, u2, wq)) { qacc = q; w = wq; }
        }
        const bool acc = qacc >= 0;
        {
            const uint32_t idx = idx0 + (acc ? 2u * (uint32_t)(qacc + 1) : 2u * NT);
            const uint32_t cb = idx >> 1;               // block holding the next draw: b0 + 1 .. b0 + NT
            uint32_t c2 = bw[NT][2], c3 = bw[NT][3];
#pragma unroll
            for (int m = NT - 1; m >= 1; m--) if (cb - b0 == (uint32_t)m) { c2 = bw[m][2]; c3 = bw[m][3]; }
            S.idx[it] = idx; S.cblock[it] = cb; S.c2[it] = c2; S.c3[it] = c3;
        }
#else
        Rng rng;
        wf_load_rng(S, it, rng);
        double eng = kinenergy<SP>(wf_get3(S, WD_P0, it));      // same p, same function as the STEP unit's test: same bits
        double B = TS.procs[sw >> 16].par[0];
        RbebConsts k = rbeb_consts(eng, B);
        double w;
        bool acc = false;
        // WF_RBEB_LOOP sequential trials per unit.  A trial accepts with p ~ 0.27, so an ionisation needs ~3.7 of them.  With
        // two trials per unit (round 1) 47 % of the entries left the unit accepted: an ionisation cost ~2 scheduling rounds
        // and ~2 evaluations of rbeb_consts (a logarithm and five divisions), and the tail below ran at 14.6 of 32 lanes —
        // the SASS page of the ncu capture attributes ~48 % of all executed instructions to ionisations.  Lanes that have
        // accepted idle while the others go on, but a trial is only ~140 instructions: measured on B200 (4e6 electrons, main
        // pass, ms) 2 trials 27.8 · 4 trials 26.4 · 5 trials 25.2 · 6 / 8 trials slower again.  Same draws, same results.
#if WF_RBEB_DYN_MIN > 0
        // up to WF_RBEB_LOOP trials, but the warp leaves the loop as soon as fewer than WF_RBEB_DYN_MIN of its lanes still need one.
        // Measured slower than the fixed five trials (24.6 ms): at most 8 trials / leave below 8 lanes 26.9, below 12 lanes 27.4,
        // at most 10 / below 5 lanes 27.3 — the ballot in the loop and the longer worst case cost more than the idle iterations.
#pragma unroll 1
        for (int q = 0; q < WF_RBEB_LOOP; q++) {
            if (!acc) {
                double u = rng.u(rc.step, rc.seed_lo, rc.seed_hi);
                double u2 = rng.u(rc.step, rc.seed_lo, rc.seed_hi);
                acc = rbeb_trial(k, u, u2, w);
            }
            const unsigned inloop = __activemask();
            if (__popc(__ballot_sync(inloop, !acc)) < WF_RBEB_DYN_MIN) break;
        }
#else
#pragma unroll 1
        for (int q = 0; q < WF_RBEB_LOOP && !acc; q++) {
            double u = rng.u(rc.step, rc.seed_lo, rc.seed_hi);
            double u2 = rng.u(rc.step, rc.seed_lo, rc.seed_hi);
            acc = rbeb_trial(k, u, u2, w);
        }
#endif
        wf_store_rng(S, it, rng);
#endif
        if (!acc) break;                  // stays an RBEB item: more trials next round
        // the accepted lanes (~3/4) finish the ionisation in this unit: one scheduling round less per event.
        // (Chaining units further -- COULOMB or IONFIN straight into the next STEP -- halves the rounds but was
        // measured 10% slower: the STEP part then runs at ~20 of 32 lanes instead of re-packed full chunks.  A separate
        // IONFIN class, one more round per ionisation, was 10 % slower too and is gone.)
        // rbeb.jl:58-80 after the sampler: kinematics, apply!(NewParticle), birth
        {
            Rng rng;
            wf_load_rng(S, it, rng);
            const double E2 = B * w;
            double E1 = eng - E2 - B;
            if (!(E2 < E1)) atomicOr(P.flags, PTL_ERR_SAMPLER_INVARIANT);   // @assert E2 < E1  rbeb.jl:63
            Vec3 p = wf_get3(S, WD_P0, it);
            Outcome o;
            const double ccut = P.pop[PTL_ELECTRON].present ? P.pop[PTL_ELECTRON].energy_cut : INFINITY;
            ionization_products(rng, rc, p, eng, E1, E2, o, ccut);
            wf_store_rng(S, it, rng);
            wf_after_collision<SP>(P, TS, S, it, o.p1, o.s1, fastsel, cut);
            if (o.sp2 >= 0) {
                uint64_t cu[2];
                child_uids(S.uid[it], rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
                long long i = S.row[it];
                add_particle(P, PTL_ELECTRON, wf_get3(S, WD_X0, it), o.p2, Q.col[COL_W][i], WFD(WD_T, it), o.s2, cu[0]);
            }
        }
        break;
    }
    // ------------------------------------------------------------------------------------------
    case WS_OTHER: {   // rare processes: the whole collide() + apply! in one unit
        // (moving this unit out of line by passing the pool by reference to a real function was measured 4.6 % SLOWER: its
        // pointers then stay live in registers across the whole kernel; see wf_other_unit for the version that does not)
        if (other_out != nullptr) nsub += other_out(&P, it, sw);
        else nsub += wf_other_unit<SP, TK>(P, Q, S, TS, fastsel, rc, cut, it, sw);
        break;
    }
    default: break;
    }
}

template <int SP, int TK, bool FIRST, bool CB>
__global__ void __launch_bounds__(WF_THREADS, WF_MIN_BLOCKS) k_advance_wf(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                              unsigned long long* row_counter, const long long* __restrict__ rows,
                                                              const unsigned long long* __restrict__ nrows) {
    if (rows != nullptr) { i0 = 0; i1 = (long long)*nrows; }     // index-list mode (rows deferred by the streaming kernel)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    // ---- carve shared memory: particle pool, then the rate table ----
    WfPool S;
    S.np = WF_THREADS;
    unsigned char* ptr = smem_raw;
    S.d = reinterpret_cast<double*>(ptr); ptr += sizeof(double) * WD_NCOL * WF_THREADS;
    S.uid = reinterpret_cast<unsigned long long*>(ptr); ptr += 8 * WF_THREADS;
    S.row = reinterpret_cast<long long*>(ptr); ptr += 8 * WF_THREADS;
    S.idx = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WF_THREADS;
    S.cblock = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WF_THREADS;
    S.c2 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WF_THREADS;
    S.c3 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WF_THREADS;
    S.state = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WF_THREADS;
    S.cnt = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * 64;
    S.order = reinterpret_cast<unsigned short*>(ptr); ptr += 2 * WF_THREADS;
    // pool bytes are a multiple of 16 by construction (see wf_pool_bytes)
    double* tsm = reinterpret_cast<double*>(smem_raw + WF_POOL_BYTES);
    // table in shared memory: Chebyshev -> cumulative rows + rate bound + process descriptors; linear -> descriptors only
    const bool fastsel = (TK == 0) && T.order == 3 && T.nprocs <= 16;
    const int nrate = (TK == 0) ? (fastsel ? WF_CUM_STRIDE * (T.k + 1) : T.order * T.nprocs * (T.k + 1)) : 0;
    const int nrb = (TK == 0) ? T.order * (T.k + 1) : 0;
    {
        const int nproc_dbl = (T.nprocs * (int)sizeof(ptl_process_desc)) / 8;
        const double* pd = reinterpret_cast<const double*>(T.procs);
        if (fastsel) {
            for (int q = threadIdx.x; q < nrate; q += blockDim.x) {
                int i = q / WF_CUM_STRIDE, rr = q - i * WF_CUM_STRIDE;
                int j = rr & 15, m = rr >> 4;
                tsm[q] = (m < 3 && j < T.nprocs) ? T.cum[m + 3 * (j + T.nprocs * i)] : (m == 0 ? WF_CUM_PAD : 0.0);
            }
        } else {
            for (int q = threadIdx.x; q < nrate; q += blockDim.x) tsm[q] = T.cum[q];
        }
        for (int q = threadIdx.x; q < nrb; q += blockDim.x) tsm[nrate + q] = T.ratebound[q];
        for (int q = threadIdx.x; q < nproc_dbl; q += blockDim.x) tsm[nrate + nrb + q] = pd[q];
    }
    const double* tcum = (TK == 0) ? tsm : T.cum;
    SmemTable TS;
    TS.rate = T.rate;                       // per-process rates stay in global memory (fallback scan only)
    TS.ratebound = tsm + nrate;
    TS.procs = reinterpret_cast<const ptl_process_desc*>(tsm + nrate + nrb);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    S.state[tid] = WS_LOAD;
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;
    __syncthreads();

    for (;;) {
        // ---- counting sort of the slots by pending work unit (warp ballots -> shared counts) ----
        const int st = (int)(S.state[tid] & 0xffu);
        int myrank = 0;
#pragma unroll
        for (int c = 0; c < WS_IDLE; c++) {
            unsigned b = __ballot_sync(0xffffffffu, st == c);
            if (lane == 0) S.cnt[c * WF_WARPS + wid] = __popc(b);
            if (st == c) myrank = __popc(b & ltmask);
        }
        __syncthreads();
        // Every warp reduces the 6 x 8 count matrix with its lanes (lane = class*8 + warp): segmented scan over
        // the warps of a class, then class totals / bases by shuffles.  ~10x fewer instructions than letting each
        // thread walk the matrix.
        int e0 = (int)S.cnt[lane];                                 // classes 0..3
        int e1 = lane < 16 ? (int)S.cnt[32 + lane] : 0;            // classes 4..5
        int s0 = e0, s1 = e1;
#pragma unroll
        for (int d = 1; d < WF_WARPS; d <<= 1) {
            int t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
            if ((lane & (WF_WARPS - 1)) >= d) { s0 += t0; s1 += t1; }
        }
        int ncls[WS_IDLE], base[WS_IDLE];
        ncls[0] = __shfl_sync(0xffffffffu, s0, 7); ncls[1] = __shfl_sync(0xffffffffu, s0, 15);
        ncls[2] = __shfl_sync(0xffffffffu, s0, 23); ncls[3] = __shfl_sync(0xffffffffu, s0, 31);
        ncls[4] = __shfl_sync(0xffffffffu, s1, 7); ncls[5] = __shfl_sync(0xffffffffu, s1, 15);
        int nwork = 0;
#pragma unroll
        for (int c = 0; c < WS_IDLE; c++) { base[c] = nwork; nwork += ncls[c]; }
        if (nwork == 0) break;                     // block-uniform: every slot is idle
        {
            int stc = st < WS_IDLE ? st : 0;
            int before0 = __shfl_sync(0xffffffffu, s0 - e0, ((stc & 3) << 3) + wid);
            int before1 = __shfl_sync(0xffffffffu, s1 - e1, ((stc & 1) << 3) + wid);
            int mybase = 0;
#pragma unroll
            for (int c = 0; c < WS_IDLE; c++) if (c == stc) mybase = base[c];
            if (st < WS_IDLE) S.order[mybase + (stc < 4 ? before0 : before1) + myrank] = (unsigned short)tid;
        }
        // Chunk scheduling: a warp executes 32 consecutive entries of ONE class, so it never mixes work units.
        // Full chunks first (class order), then the largest partial chunks; smaller remainders wait for a
        // later round (they only grow), which keeps ~90 % of the lanes busy with zero intra-warp divergence.
        int my_c = -1, my_k = 0;
        {
            int nfull = 0;
#pragma unroll
            for (int c = 0; c < WS_IDLE; c++) {
                int f = ncls[c] >> 5;
                if (my_c < 0 && wid < nfull + f) { my_c = c; my_k = wid - nfull; }
                nfull += f;
            }
            if (my_c < 0) {                        // (wid - nfull)-th largest remainder, ranked across lanes 0..5
                int want = wid - nfull;
                int key = 0;
#pragma unroll
                for (int c = 0; c < WS_IDLE; c++) if (lane == c) key = (ncls[c] & 31) ? (((ncls[c] & 31) << 3) | (WS_IDLE - 1 - c)) : 0;
                int rank = 0;
#pragma unroll
                for (int c = 0; c < WS_IDLE; c++) rank += __shfl_sync(0xffffffffu, key, c) > key;
                unsigned pick = __ballot_sync(0xffffffffu, lane < WS_IDLE && key > 0 && rank == want);
                if (pick) {
                    my_c = __ffs(pick) - 1;
#pragma unroll
                    for (int c = 0; c < WS_IDLE; c++) if (c == my_c) my_k = ncls[c] >> 5;
                }
            }
        }
        __syncthreads();

        int pos = my_k * 32 + lane;
        bool has = false;
        int it = 0;
        if (my_c >= 0) {
            int nc = 0, bc = 0;
#pragma unroll
            for (int c = 0; c < WS_IDLE; c++) if (c == my_c) { nc = ncls[c]; bc = base[c]; }
            has = pos < nc;
            if (has) it = (int)S.order[bc + pos];
        }
        const uint32_t sw = has ? S.state[it] : (uint32_t)WS_IDLE;
        const int cls = (int)(sw & 0xffu);
        const unsigned ldmask = __ballot_sync(0xffffffffu, has && cls == WS_LOAD);

        if (has) wf_execute_unit<SP, TK, FIRST, CB>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, ldmask, lane, ltmask, row_counter, i0, i1, nsub, rows);
        __syncthreads();
    }

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
}

}  // namespace ptl
