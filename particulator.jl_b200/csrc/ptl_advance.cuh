// ptl_advance.cuh — K1: the fused advance kernel (one specialisation per species).
//
// Replaces, for one population and one pass of advance1! (mixed_population.jl:56-93):
//   advance_init!/setr!  (mixed_population.jl:97-110, collisions.jl:63-74)      [FIRST pass only]
//   the per-particle sub-step loop (mixed_population.jl:61-88)
//   advance_particle (pusher.jl:41-76), onadvance of WallCallback (callback.jl:167-184)
//   do_one_collision! (collisions.jl:142-199), collide (9 processes + LXCat kinds), apply! (:83-133)
//   add_particle! (population.jl:103-113) with warp-aggregated atomic appends.
//
// Mapping: persistent warps; each warp claims tiles of 32 consecutive rows from a global tile
// counter (dynamic load balance at warp granularity — the sub-step count per particle varies by
// >10x with energy), one particle per lane, state in registers for the whole dt, every column
// access of a warp is one contiguous 256-byte span.  The Chebyshev rate table of the species
// (<= 12 KB) is staged once per block in shared memory.
#pragma once
#include "ptl_physics.cuh"

namespace ptl {

constexpr int ADV_THREADS = 128;

struct SmemTable {
    const double* rate;        // [order, nprocs, k+1] (cheb) in shared memory, or global pointer (linear)
    const double* ratebound;   // [order, k+1]
    const ptl_process_desc* procs;
};

// add_particle!(popl, state): population.jl:103-113 + setr! of the newborn (collisions.jl:104,118,128,132)
static __device__ __noinline__ void add_particle(const AdvanceParams& P, int sp, Vec3 x, Vec3 p, double w, double t, double s, uint64_t uid) {
    const PopView& Q = P.pop[sp];
    if (!Q.present) return;
    double eng = kinenergy_rt(sp, p);
    if (eng <= Q.energy_cut) return;
    // warp-aggregated append: lanes of the converged group that target the same population share one atomic
    unsigned grp = __match_any_sync(__activemask(), sp);
    int leader = __ffs(grp) - 1;
    int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(Q.n, (unsigned long long)__popc(grp));
    base = __shfl_sync(grp, base, leader);
    long long slot = (long long)base + __popc(grp & ((1u << lane) - 1));
    if (lane == leader) atomicAdd(P.births, (unsigned long long)__popc(grp));
    if (slot >= Q.capacity) {   // @assert n < length(particles)  population.jl:107
        atomicOr(P.flags, PTL_ERR_CAPACITY_OVERFLOW);
        return;
    }
    Q.col[COL_X0][slot] = x.x; Q.col[COL_X1][slot] = x.y; Q.col[COL_X2][slot] = x.z;
    Q.col[COL_P0][slot] = p.x; Q.col[COL_P1][slot] = p.y; Q.col[COL_P2][slot] = p.z;
    Q.col[COL_W][slot] = w; Q.col[COL_T][slot] = t; Q.col[COL_S][slot] = s;
    Q.col[COL_R][slot] = ratebound_global(P.tab[sp], eng, P.flags);
    Q.active[slot] = 1;
    Q.uid[slot] = uid;
}

template <int SP>
__device__ __forceinline__ double own_ratebound(const AdvanceParams& P, const SmemTable& S, double eng) {
    const TableView& T = P.tab[SP];
    if (T.kind == 0) {
        Pre pre = precheb(eng, T.k, T.xmax, T.rxmax);
        if (pre.oob) atomicOr(P.flags, PTL_ERR_ENERGY_OUT_OF_TABLE);
        return chebsum(S.ratebound + T.order * pre.i, pre, T.order);
    }
    return ratebound_global(T, eng, P.flags);
}

// setr!: collisions.jl:63-74
template <int SP>
static __device__ __noinline__ double setr(const AdvanceParams& P, const SmemTable& S, Vec3 p) {
    double eng = kinenergy<SP>(p);
    if (eng < P.pop[SP].energy_cut) return 0.0;
    return own_ratebound<SP>(P, S, eng);
}

// setr! for the common table shape (Chebyshev, order 3), inline on the shared-memory rate-bound rows: ~55 instructions.
// The generic setr<SP> is an out-of-line call that reads AdvanceParams through a generic pointer and carries the
// any-order loop and the IEEE-division fallback (112 executed instructions per call, 6 % of the kernel in ncu).
// (Measured: as a real function, -DPTL_SETR3_NOINLINE, 29.7 ms against 27.6 ms inline for the main pass of 4e6 electrons.)
template <int SP>
#ifdef PTL_SETR3_NOINLINE
__device__ __noinline__ double wf_setr_cheb3(
#else
__device__ __forceinline__ double wf_setr_cheb3(
#endif
const AdvanceParams& P, const TableView& T, const double* __restrict__ rb, double cut, Vec3 p) {
    const double eng = kinenergy<SP>(p);
    if (eng < cut) return 0.0;                                                         // collisions.jl:66
    const Pre pre = precheb(eng, T.k, T.xmax, T.rxmax);
    if (pre.oob) atomicOr(P.flags, PTL_ERR_ENERGY_OUT_OF_TABLE);
    const double* a = rb + 3 * pre.i;
    return __dadd_rn(__dadd_rn(a[0], __dmul_rn(a[1], pre.a)), __dmul_rn(a[2], pre.b));   // chebsum, order 3
}

// Repeated below-cut sub-steps taken in blocks (used by every advance kernel; see the comment in ptl_advance_wf.cuh).
// A particle that the field decelerated below energy_cut keeps its old r != 0: do_one_collision! returns before it draws
// anything (collisions.jl:148-151), so the loop of mixed_population.jl:66-87 repeats the same dt = s/r push until tfinal or
// until the energy is back above the cut.  Under the uniform-E fast path the kinetic energy is a convex function of time,
// so "below the cut after j sub-steps" implies below the cut at every sub-step in between: take the j sub-steps as ONE RK2
// step of j*dt (p exact up to rounding, x within the RK2 truncation error of the longer step, < 1e-10 relative), with t
// and trem accumulated sub-step by sub-step so that they stay bit-identical, and halve j when the energy would cross
// the cut so that the first test at or above the cut is still taken by the regular path on the reference's own time grid.
// Takes at most ~max_work sub-steps per call; `more` tells whether further repeated sub-steps remain.
template <int SP>
__device__ __forceinline__ unsigned long long coast_below_cut(const AdvanceParams& P, Vec3& x, Vec3& p, double& t, double& trem, const double tnext,
                                                              const double cut, const int max_work, bool& more) {
    unsigned long long nsub = 0;
    int M = 128, work = 0;
    more = false;
    while (M >= 1) {
        double tr = trem, tt = t;
        int j = 0;
        for (; j < M && tr > DBL_EPS && tr > tnext; j++) { tr -= tnext; tt += tnext; }   // mixed_population.jl:66-68,86
        if (j == 0) break;                                           // the final free flight belongs to the regular path
        work += j;
        Vec3 xc = x, pc = p;
        double tdum = 0;
        push<SP>(P, xc, pc, tdum, tnext * j);
        if (!(kinenergy<SP>(pc) < cut)) { M = j >> 1; continue; }    // would reach the cut: halve (M == 0: the regular path takes it)
        x = xc; p = pc; t = tt; trem = tr; nsub += (unsigned long long)j;
        if (work >= max_work) { more = true; break; }
    }
    return nsub;
}

// `rows_per_warp` (1..32): how many lanes of a warp carry a particle.  A lane-per-particle warp executes the UNION of its lanes'
// control flow, so one loop iteration costs push + every process branch some lane takes; in the latency regime (a pass of a few
// thousand rows, bound by the ~400-collision chain of its slowest particle) giving every particle a warp of its own halves the
// time per sub-step.  The launcher picks the smallest value that still fills the machine.
template <int SP, bool FIRST, bool CB>
__global__ void __launch_bounds__(ADV_THREADS) k_advance(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                         unsigned long long* tile_counter, const long long* __restrict__ rows,
                                                         const unsigned long long* __restrict__ nrows, const int rows_per_warp) {
    // index-list mode: process rows[0 .. *nrows) (the particles the streaming kernel deferred) instead of [i0, i1)
    if (rows != nullptr) { i0 = 0; i1 = (long long)*nrows; }
    extern __shared__ double smem[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    SmemTable S;
    if (T.kind == 0) {
        int nrate = T.order * T.nprocs * (T.k + 1), nrb = T.order * (T.k + 1);
        int nproc_dbl = (T.nprocs * (int)sizeof(ptl_process_desc)) / 8;
        for (int q = threadIdx.x; q < nrate; q += blockDim.x) smem[q] = T.rate[q];
        for (int q = threadIdx.x; q < nrb; q += blockDim.x) smem[nrate + q] = T.ratebound[q];
        const double* pd = reinterpret_cast<const double*>(T.procs);
        for (int q = threadIdx.x; q < nproc_dbl; q += blockDim.x) smem[nrate + nrb + q] = pd[q];
        S.rate = smem;
        S.ratebound = smem + nrate;
        S.procs = reinterpret_cast<const ptl_process_desc*>(smem + nrate + nrb);
    } else {
        int nproc_dbl = (T.nprocs * (int)sizeof(ptl_process_desc)) / 8;
        const double* pd = reinterpret_cast<const double*>(T.procs);
        for (int q = threadIdx.x; q < nproc_dbl; q += blockDim.x) smem[q] = pd[q];
        S.rate = T.rate;
        S.ratebound = nullptr;
        S.procs = reinterpret_cast<const ptl_process_desc*>(smem);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;

    for (;;) {
        long long tile = 0;
        if (lane == 0) tile = (long long)atomicAdd(tile_counter, 1ULL);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        long long base = i0 + tile * rows_per_warp;
        if (base >= i1) break;
        long long i = base + lane;
        if (lane >= rows_per_warp || i >= i1) continue;
        if (rows != nullptr) i = rows[i];
        if (!Q.active[i]) continue;    // l.active || continue   mixed_population.jl:63

        Vec3 x = {Q.col[COL_X0][i], Q.col[COL_X1][i], Q.col[COL_X2][i]};
        Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
        double w = Q.col[COL_W][i], t = Q.col[COL_T][i], s = Q.col[COL_S][i];
        double r = FIRST ? setr<SP>(P, S, p) : Q.col[COL_R][i];   // advance_init!  mixed_population.jl:97-110
        uint64_t uid = Q.uid[i];
        Rng rng;
        rng.init(uid, DOM_COLLISION);
        bool act = true;

        double trem = P.tfinal - t;                         // :65
        while (trem > DBL_EPS && act) {                     // :66
            double tnext = s / r;                           // :67  (r == 0 -> Inf -> free flight)
            bool collides = trem > tnext;                   // :68
            double dt = collides ? tnext : trem;
            if (!collides) s -= dt * r;                     // :74
            Vec3 xo = x, po = p;
            double wo = w, to = t;
            push<SP>(P, x, p, t, dt);                       // :77
            if (CB) {                                       // onadvance(WallCallback)  callback.jl:167-184
                for (int k = 0; k < P.cb.nwalls; k++) {
                    const ptl_wall_desc& wd = P.cb.wall[k];
                    if (wd.species != SP) continue;
                    double xoc = wd.coord == 0 ? xo.x : (wd.coord == 1 ? xo.y : xo.z);
                    double xnc = wd.coord == 0 ? x.x : (wd.coord == 1 ? x.y : x.z);
                    if (xoc < wd.v && wd.v < xnc) {
                        double f = (wd.v - xoc) / (xnc - xoc);
                        rng.skip();   // lincomb builds a state with the 4-arg ctor: one discarded s draw (electron.jl:127-132)
                        const WallBuf& W = P.wall[k];
                        unsigned long long slot = atomicAdd(W.n, 1ULL);
                        if ((long long)slot < W.capacity) {
                            W.col[0][slot] = x.x * f + xo.x * (1 - f); W.col[1][slot] = x.y * f + xo.y * (1 - f); W.col[2][slot] = x.z * f + xo.z * (1 - f);
                            W.col[3][slot] = p.x * f + po.x * (1 - f); W.col[4][slot] = p.y * f + po.y * (1 - f); W.col[5][slot] = p.z * f + po.z * (1 - f);
                            W.col[6][slot] = w * f + wo * (1 - f);
                            W.col[7][slot] = t * f + to * (1 - f);
                        } else {
                            atomicOr(P.flags, PTL_ERR_CAPACITY_OVERFLOW);
                        }
                        if (wd.drop) act = false;
                    }
                }
            }
            if (collides && act) {                          // :83  do_one_collision!  collisions.jl:142-199
                double eng;
                if (r != 0.0 && (eng = kinenergy<SP>(p)) >= cut) {   // :148-151
                    Pre pre = (T.kind == 0) ? precheb(eng, T.k, T.xmax, T.rxmax) : indweight(T, eng);   // :153
                    if (pre.oob) atomicOr(P.flags, PTL_ERR_ENERGY_OUT_OF_TABLE);
                    rng.idx += rng.idx & 1u;   // every collision test starts on an even draw index (Philox block boundary)
                    double xi = rng.u(rc.step, rc.seed_lo, rc.seed_hi) * r;                     // :154
                    int jsel = -1;
                    const int np = T.nprocs;
                    for (int j = 0; j < np; j++) {          // :166-180
                        double nu = (T.kind == 0) ? chebsum(S.rate + T.order * (j + np * pre.i), pre, T.order)
                                                  : linear_rate(S.rate, np, j, pre);
                        if (nu > xi) { jsel = j; break; }
                        xi -= nu;
                    }
                    Outcome o;
                    o.kind = OUT_NULL;
                    if (jsel >= 0) {
                        collide<SP>(rng, rc, P, S.procs[jsel], p, eng, o);
                    } else if (!(xi >= 0)) {
                        atomicOr(P.flags, PTL_ERR_RATE_BOUND_VIOLATED);     // :186
                    }
                    if (CB && P.cb.count_collisions) atomicAdd(T.counts + (jsel >= 0 ? jsel : np), 1ULL);
                    // apply!: collisions.jl:83-133
                    switch (o.kind) {
                    case OUT_NULL:
                        r = setr<SP>(P, S, p);
                        s = -nlog(rng.u(rc.step, rc.seed_lo, rc.seed_hi));
                        break;
                    case OUT_STATE_CHANGE:
                        p = o.p1; s = o.s1; r = setr<SP>(P, S, p);
                        break;
                    case OUT_NEW_PARTICLE: {
                        p = o.p1; s = o.s1; r = setr<SP>(P, S, p);
                        uint64_t cu[2];
                        child_uids(uid, rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
                        add_particle(P, o.sp2, x, o.p2, w, t, o.s2, cu[0]);
                        break;
                    }
                    case OUT_REMOVE: act = false; break;
                    case OUT_REPLACE: {
                        act = false;
                        uint64_t cu[2];
                        child_uids(uid, rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
                        add_particle(P, o.sp2, x, o.p2, w, t, o.s2, cu[0]);
                        break;
                    }
                    case OUT_REPLACE_PAIR: {
                        act = false;
                        uint64_t cu[2];
                        child_uids(uid, rng.idx, rc.step, rc.seed_lo, rc.seed_hi, cu);
                        add_particle(P, o.sp2, x, o.p2, w, t, o.s2, cu[0]);
                        add_particle(P, o.sp3, x, o.p3, w, t, o.s3, cu[1]);
                        break;
                    }
                    }
                } else if (!CB && SP != PTL_PHOTON && r != 0.0 && P.fast_force && trem - dt > tnext) {
                    // below the cut with r != 0: the loop would repeat this same dt = s/r push; take the repeats in blocks
                    trem -= dt;                              // :86 of the sub-step just taken
                    nsub++;
                    bool more;
                    do { nsub += coast_below_cut<SP>(P, x, p, t, trem, tnext, cut, 1 << 20, more); } while (more);
                    continue;
                }
            }
            trem -= dt;                                      // :86
            nsub++;
        }

        Q.col[COL_X0][i] = x.x; Q.col[COL_X1][i] = x.y; Q.col[COL_X2][i] = x.z;
        Q.col[COL_P0][i] = p.x; Q.col[COL_P1][i] = p.y; Q.col[COL_P2][i] = p.z;
        Q.col[COL_T][i] = t; Q.col[COL_S][i] = s; Q.col[COL_R][i] = r;
        if (!act) Q.active[i] = 0;
    }

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
}

// K1s: streaming fast path for species with kappa << 1 (photons: mean free path of metres vs c*dt = 7.5 mm).
// A particle whose next event lies beyond tfinal does exactly ONE sub-step of the reference loop
// (mixed_population.jl:66-87 with collides == false): s -= trem*r, one push, t += trem.  That is pure streaming:
// two particles per thread, 128-bit loads of nine columns (+ active), stores of only the columns that change
// (x, t, s; r when advance_init! changed it).  Particles that do collide within dt (~1 %) are left untouched and
// appended to an index list which the general kernel processes afterwards.
constexpr int STREAM_THREADS = 256;
#ifndef STREAM_LEPTON_MINB
#define STREAM_LEPTON_MINB 2              // resident CTAs per SM the lepton streaming kernel is compiled for (register budget 65536 / 256 / this).
                                          // Measured on B200, 1e7 electrons at kappa ~ 1: 2 CTAs (128 registers, no spills) 0.385 ms = 64 % of the measured
                                          // HBM bandwidth; 3 CTAs (80 registers, 18 % of the instructions are spill traffic) 0.461 ms = 54 %; 4 CTAs 0.514 ms
#endif

template <int SP, bool FIRST>
// (No minimum-blocks bound on purpose: ptxas picks 118 registers, 2 CTAs per SM, 0.39 ms for 2e7 photons = 87 % of the
// measured HBM bandwidth.  Forcing 3 or 4 CTAs (80 / 64 registers, spills) measured 0.41 / 0.43 ms; (256, 1) 0.61 ms.)
// Leptons (species that a force acts on) carry the RK2 push and the rate-bound lookup: two of them per thread took 216
// registers, ONE 256-thread CTA per SM, and the kernel could not cover the HBM latency (48 % of the measured bandwidth at
// kappa ~ 1).  They go one particle per thread (STREAM_NP = 1: 64-bit accesses, still one full 256-byte span per warp and
// column) at three CTAs per SM; photons keep two per thread and 128-bit accesses.
__global__ void __launch_bounds__(STREAM_THREADS, SP == PTL_PHOTON ? 0 : STREAM_LEPTON_MINB) k_advance_stream(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                                  long long* __restrict__ slow_rows, unsigned long long* slow_count) {
    extern __shared__ double smem[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    SmemTable S;
    S.rate = nullptr; S.procs = nullptr;
    if (T.kind == 0) {
        int nrb = T.order * (T.k + 1);
        for (int q = threadIdx.x; q < nrb; q += blockDim.x) smem[q] = T.ratebound[q];
        S.ratebound = smem;
    } else {
        S.ratebound = nullptr;
    }
    __syncthreads();
    unsigned long long nsub = 0;
    // the usual table shape: setr! inline on the shared-memory rate bound (~55 instructions).  Leptons only: two inline copies
    // took the photon kernel (two particles per thread) from 118 to 157 registers, one CTA per SM, 0.39 -> 0.52 ms for 2e7 photons
    const bool cheb3 = SP != PTL_PHOTON && T.kind == 0 && T.order == 3;
    const double cut = Q.energy_cut;
    constexpr int NP = (SP == PTL_PHOTON) ? 2 : 1;      // particles per thread
    const long long npairs = (i1 - i0 + NP - 1) / NP;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < npairs; q += (long long)gridDim.x * blockDim.x) {
        const long long i = i0 + NP * q;         // i0 is even (rows of a pass start at 0 or at an even boundary, see launcher)
        const bool two = NP == 2 && i + 1 < i1;
        double2 c[9];
        unsigned char a0, a1 = 0;
        if (two) {
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const int col = k < 6 ? k : k + 1;      // x0..p2, t, s, r  (w is not needed)
                c[k] = *reinterpret_cast<const double2*>(Q.col[col] + i);
            }
            uchar2 aa = *reinterpret_cast<const uchar2*>(Q.active + i);
            a0 = aa.x; a1 = aa.y;
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const int col = k < 6 ? k : k + 1;
                c[k] = make_double2(Q.col[col][i], 0.0);
            }
            a0 = Q.active[i];
        }
        double xs[2][3], ps[2][3], ts[2], ss[2], rs[2];
        bool wr[2] = {false, false}, wr_r[2] = {false, false};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const bool act = h == 0 ? (a0 != 0) : (two && a1 != 0);
            Vec3 x = {h ? c[0].y : c[0].x, h ? c[1].y : c[1].x, h ? c[2].y : c[2].x};
            Vec3 p = {h ? c[3].y : c[3].x, h ? c[4].y : c[4].x, h ? c[5].y : c[5].x};
            double t = h ? c[6].y : c[6].x, s = h ? c[7].y : c[7].x, r0 = h ? c[8].y : c[8].x;
            if (!act) continue;
            double r = FIRST ? (cheb3 ? wf_setr_cheb3<SP>(P, T, S.ratebound, cut, p) : setr<SP>(P, S, p)) : r0;           // advance_init!
            double trem = P.tfinal - t;
            if (!(trem > DBL_EPS)) {                              // nothing to do this step (:66)
                if (FIRST && r != r0) { wr_r[h] = true; rs[h] = r; }
                continue;
            }
            double tnext = s / r;
            if (trem > tnext) {                                   // collides within dt: defer to the general kernel
                unsigned long long k = atomicAdd(slow_count, 1ULL);
                slow_rows[k] = i + h;
                continue;
            }
            s -= trem * r;                                        // :74
            push<SP>(P, x, p, t, trem);                           // :77 (photons: p unchanged)
            nsub++;
            xs[h][0] = x.x; xs[h][1] = x.y; xs[h][2] = x.z; ts[h] = t; ss[h] = s; rs[h] = r;
            wr[h] = true;
            wr_r[h] = FIRST && r != r0;
            if (SP != PTL_PHOTON) { ps[h][0] = p.x; ps[h][1] = p.y; ps[h][2] = p.z; }   // species whose momentum changes under the pusher
        }
        if (two && wr[0] && wr[1]) {
            *reinterpret_cast<double2*>(Q.col[COL_X0] + i) = make_double2(xs[0][0], xs[1][0]);
            *reinterpret_cast<double2*>(Q.col[COL_X1] + i) = make_double2(xs[0][1], xs[1][1]);
            *reinterpret_cast<double2*>(Q.col[COL_X2] + i) = make_double2(xs[0][2], xs[1][2]);
            *reinterpret_cast<double2*>(Q.col[COL_T] + i) = make_double2(ts[0], ts[1]);
            *reinterpret_cast<double2*>(Q.col[COL_S] + i) = make_double2(ss[0], ss[1]);
            if (SP != PTL_PHOTON) {
                *reinterpret_cast<double2*>(Q.col[COL_P0] + i) = make_double2(ps[0][0], ps[1][0]);
                *reinterpret_cast<double2*>(Q.col[COL_P1] + i) = make_double2(ps[0][1], ps[1][1]);
                *reinterpret_cast<double2*>(Q.col[COL_P2] + i) = make_double2(ps[0][2], ps[1][2]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (!wr[h]) continue;
                Q.col[COL_X0][i + h] = xs[h][0]; Q.col[COL_X1][i + h] = xs[h][1]; Q.col[COL_X2][i + h] = xs[h][2];
                Q.col[COL_T][i + h] = ts[h]; Q.col[COL_S][i + h] = ss[h];
                if (SP != PTL_PHOTON) { Q.col[COL_P0][i + h] = ps[h][0]; Q.col[COL_P1][i + h] = ps[h][1]; Q.col[COL_P2][i + h] = ps[h][2]; }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) if (wr_r[h]) Q.col[COL_R][i + h] = rs[h];
    }
    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if ((threadIdx.x & 31) == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
}

// K1t: the streaming fast path for LEPTONS with the rows staged through shared memory by the TMA engine.
// k_advance_stream<lepton> (one row per thread, plain loads) reaches 54 % of the measured HBM bandwidth at kappa ~ 1
// (profiles/r2_electron_stream_kernel_ncu_summary.csv): 24 resident warps per SM each issue ten loads and then spend ~400
// instructions on setr! and the RK2 push before they ask for memory again — long-scoreboard stall 12.9 warps per issue, the
// memory system idles while the warps compute.  Here the loads do not belong to the warps: a persistent CTA walks over tiles
// of STT_ROWS rows; a producer warp arms an mbarrier and issues ten bulk copies (cp.async.bulk, nine double columns + the
// active bytes) per tile into a ring of STT_STAGES shared-memory stages, STT_STAGES - 1 tiles ahead of the one being
// computed, so ~50 KB per CTA are in flight whatever the warps are doing.  Results go back with plain coalesced stores
// (fire and forget).  Same arithmetic, same deferral of rows that collide within dt as k_advance_stream.
constexpr int STT_ROWS = 256;            // rows per tile = threads per CTA
#ifndef STT_STAGES_MACRO
#define STT_STAGES_MACRO 4
#endif
#ifndef STT_MINB
#define STT_MINB 2                        // resident CTAs per SM (3 CTAs = 72 registers: ncu showed the spills missing the small L1 that 170 KB of shared memory leave)
#endif
constexpr int STT_STAGES = STT_STAGES_MACRO;
constexpr int STT_STAGE_BYTES = 9 * STT_ROWS * 8 + STT_ROWS;     // nine double columns + active bytes
constexpr size_t STT_RING_BYTES = (size_t)STT_STAGES * STT_STAGE_BYTES + 16 * STT_STAGES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// STT_ROWS consumer threads (one row each per tile) + one producer warp.  Two mbarriers per stage: `full` (the producer's
// expect_tx arrival + the bytes of the ten bulk copies) and `empty` (one arrival per consumer warp once its 32 rows are
// read), so no CTA-wide barrier: a consumer warp that finds its next tile landed goes on at once.  (The first version
// refilled from thread 0 behind a __syncthreads per tile and was slower than the plain kernel: 0.53 against 0.45 ms for
// 1e7 electrons — every warp waited for the slowest row of the tile.)
constexpr int STT_THREADS = STT_ROWS + 32;
template <int SP, bool FIRST>
__global__ void __launch_bounds__(STT_THREADS, STT_MINB) k_advance_stream_tma(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                                      long long* __restrict__ slow_rows, unsigned long long* slow_count) {
    extern __shared__ __align__(128) unsigned char stt_smem[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stt_smem + (size_t)STT_STAGES * STT_STAGE_BYTES);   // full[S], empty[S]
    double* rbs = reinterpret_cast<double*>(stt_smem + STT_RING_BYTES);
    SmemTable S;
    S.rate = nullptr; S.procs = nullptr;
    if (T.kind == 0) {
        const int nrb = T.order * (T.k + 1);
        for (int q = threadIdx.x; q < nrb; q += blockDim.x) rbs[q] = T.ratebound[q];
        S.ratebound = rbs;
    } else {
        S.ratebound = nullptr;
    }
    if (threadIdx.x == 0) {
        for (int st = 0; st < STT_STAGES; st++) { mbar_init(smem_u32(bars + st), 1); mbar_init(smem_u32(bars + STT_STAGES + st), STT_ROWS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const bool cheb3 = T.kind == 0 && T.order == 3;
    const double cut = Q.energy_cut;
    const long long ntiles = (i1 - i0 + STT_ROWS - 1) / STT_ROWS;
    unsigned long long nsub = 0;
    if (threadIdx.x >= STT_ROWS) {
        // ---- producer warp: one lane walks over this CTA's tiles and keeps the ring full ----
        if (threadIdx.x == STT_ROWS) {
            int st = 0;
            uint32_t parity = 1;        // parity of the `empty` phase to wait for; the first pass over the ring waits for nothing
            bool wrapped = false;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (wrapped) mbar_wait(smem_u32(bars + STT_STAGES + st), parity);
                // i0 is a multiple of 16 (launcher): every tile starts on a 16-byte boundary of every column, the byte column included
                const long long r0 = i0 + tile * STT_ROWS;
                long long nr = i1 - r0;
                if (nr > STT_ROWS) nr = STT_ROWS;
                const uint32_t nb8 = (uint32_t)((nr + 1) & ~1LL) * 8u;      // bulk copies move multiples of 16 bytes: the padding of a
                const uint32_t nb1 = (uint32_t)((nr + 15) & ~15LL);         // partial last tile lies inside the 256-byte aligned column
                const uint32_t bar = smem_u32(bars + st);
                unsigned char* base = stt_smem + (size_t)st * STT_STAGE_BYTES;
                mbar_expect_tx(bar, 9u * nb8 + nb1);
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int col = k < 6 ? k : k + 1;                      // x0..p2, t, s, r  (w is not needed)
                    bulk_g2s(smem_u32(base + (size_t)k * STT_ROWS * 8), Q.col[col] + r0, nb8, bar);
                }
                bulk_g2s(smem_u32(base + 9 * STT_ROWS * 8), Q.active + r0, nb1, bar);
                if (++st == STT_STAGES) { st = 0; parity = wrapped ? (parity ^ 1u) : 0u; wrapped = true; }
            }
        }
        return;
    }
    // ---- consumer warps ----
    int st = 0;
    uint32_t parity = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(smem_u32(bars + st), parity);
        const unsigned char* base = stt_smem + (size_t)st * STT_STAGE_BYTES;
        const double* cd = reinterpret_cast<const double*>(base);
        const long long i = i0 + tile * STT_ROWS + threadIdx.x;
        const int q = threadIdx.x;
        const bool live = i < i1 && base[9 * STT_ROWS * 8 + q] != 0;
        Vec3 x = {cd[q], cd[STT_ROWS + q], cd[2 * STT_ROWS + q]};
        Vec3 p = {cd[3 * STT_ROWS + q], cd[4 * STT_ROWS + q], cd[5 * STT_ROWS + q]};
        double t = cd[6 * STT_ROWS + q], s = cd[7 * STT_ROWS + q];
        const double r0 = cd[8 * STT_ROWS + q];
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(bars + STT_STAGES + st));     // this warp's rows are in registers
        if (live) {
            const double r = FIRST ? (cheb3 ? wf_setr_cheb3<SP>(P, T, S.ratebound, cut, p) : setr<SP>(P, S, p)) : r0;          // advance_init!
            const double trem = P.tfinal - t;
            if (!(trem > DBL_EPS)) {                                  // nothing to do this step (:66)
                if (FIRST && r != r0) Q.col[COL_R][i] = r;
            } else if (trem > s / r) {                                // collides within dt: defer to the general kernel
                unsigned long long kk = atomicAdd(slow_count, 1ULL);
                slow_rows[kk] = i;
            } else {
                s -= trem * r;                                        // :74
                push<SP>(P, x, p, t, trem);                           // :77
                nsub++;
                Q.col[COL_X0][i] = x.x; Q.col[COL_X1][i] = x.y; Q.col[COL_X2][i] = x.z;
                Q.col[COL_P0][i] = p.x; Q.col[COL_P1][i] = p.y; Q.col[COL_P2][i] = p.z;
                Q.col[COL_T][i] = t; Q.col[COL_S][i] = s;
                if (FIRST && r != r0) Q.col[COL_R][i] = r;
            }
        }
        if (++st == STT_STAGES) { st = 0; parity ^= 1u; }
    }
    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if ((threadIdx.x & 31) == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
}

// init!(mpopl) / advance_init!: setr! on all actives (mixed_population.jl:20-35)
template <int SP>
__global__ void k_init_r(const __grid_constant__ AdvanceParams P, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PopView& Q = P.pop[SP];
    if (!Q.active[i]) return;
    Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
    double eng = kinenergy<SP>(p);
    Q.col[COL_R][i] = eng < Q.energy_cut ? 0.0 : ratebound_global(P.tab[SP], eng, P.flags);
}

// bit-exact tier entry point: presample + rate(j) + ratebound for n energies
static __global__ void k_table_eval(TableView T, long long n, const double* __restrict__ energy, double* __restrict__ rates,
                             double* __restrict__ bound, int* flags) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double eng = energy[i];
    Pre pre = (T.kind == 0) ? precheb(eng, T.k, T.xmax, T.rxmax) : indweight(T, eng);
    for (int j = 0; j < T.nprocs; j++)
        rates[j + (size_t)T.nprocs * i] = (T.kind == 0) ? chebsum(T.rate + (size_t)T.order * (j + (size_t)T.nprocs * pre.i), pre, T.order)
                                                        : linear_rate(T.rate, T.nprocs, j, pre);
    bound[i] = ratebound_global(T, eng, flags);
}

// deterministic replay of single collide() events (test entry point ptl_collide_test)
template <int SP>
__global__ void k_collide_test(const __grid_constant__ AdvanceParams P, TableView T, int j, long long n, const double* __restrict__ p3,
                               unsigned long long uid0, double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    Rng rng;
    rng.init(uid0 + (unsigned long long)i, DOM_COLLISION);
    Vec3 p = {p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]};
    Outcome o;
    o.kind = OUT_NULL; o.sp2 = o.sp3 = 0;
    o.p1 = o.p2 = o.p3 = {0, 0, 0};
    o.s1 = o.s2 = o.s3 = 0;
    collide<SP>(rng, rc, P, T.procs[j], p, kinenergy<SP>(p), o);
    double* r = out + 24 * i;
    for (int q = 0; q < 24; q++) r[q] = 0;
    r[0] = o.kind; r[3] = rng.idx;
    if (o.kind == OUT_STATE_CHANGE || o.kind == OUT_NEW_PARTICLE) { r[4] = o.p1.x; r[5] = o.p1.y; r[6] = o.p1.z; r[7] = o.s1; }
    if (o.kind == OUT_NEW_PARTICLE || o.kind == OUT_REPLACE || o.kind == OUT_REPLACE_PAIR) { r[1] = o.sp2; r[8] = o.p2.x; r[9] = o.p2.y; r[10] = o.p2.z; r[11] = o.s2; }
    if (o.kind == OUT_REPLACE_PAIR) { r[2] = o.sp3; r[12] = o.p3.x; r[13] = o.p3.y; r[14] = o.p3.z; r[15] = o.s3; }
}

}  // namespace ptl
