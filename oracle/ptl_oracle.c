/*
 * ptl_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the particle-advance hot path of aluque/Particulator.jl,
 * written by following the reference source function by function (citations are
 * `file:line` relative to the reference tree).  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the reference ships an empty test suite (test/runtests.jl:4-6), no
 * golden vectors, and Julia is not installable in this image, so this restatement could
 * not be checked against outputs of the reference itself.  It is pinned instead by
 * analytic known answers of the physics the reference cites (tests/test_oracle_*.py).
 *
 * Differences from the reference that are deliberate and documented in DESIGN.md:
 *   - rand() (task-local Xoshiro, src/util.jl:17) is replaced by a counter-based
 *     Philox4x32-10 stream per particle uid, uniform in the OPEN interval (0,1);
 *   - turn() guards its two NaN poles (src/util.jl:40-57);
 *   - @assert failures become sticky error bits instead of exceptions;
 *   - the stale slow-electron state (src/slow-electron.jl:9-17) is read consistently
 *     with the current advance loop.
 *
 * Floating point: compiled with -ffp-contract=off so that, like Julia, a*b+c is two
 * roundings.  Expression order follows the Julia source.
 *
 * The exported functions mirror include/particulator_b200.h with the prefix ora_.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/particulator_b200.h"

/* ------------------------------------------------------------------------------------ */
/* constants: src/constants.jl (CODATA-2014 via scipy)                                   */
/* ------------------------------------------------------------------------------------ */
#define CO_C 299792458.0                   /* constants.jl:37  */
#define CO_E 1.6021766208e-19              /* constants.jl:50-54 (e = eV = elementary_charge) */
#define CO_ME 9.10938356e-31               /* constants.jl:52  */
#define CO_EPS0 8.854187817620389e-12      /* constants.jl:55  */
#define CO_ALPHA 0.0072973525664           /* constants.jl:60  */
#define CO_HBAR 1.0545718001391127e-34     /* constants.jl:77  */
#define CO_PI 3.141592653589793

static double R_E, A_0, MC2, MC; /* constants.jl:157-164, computed in the same expression order */

static void init_constants(void) {
    R_E = (CO_E * CO_E) / (CO_ME * (CO_C * CO_C)) / (4 * CO_PI * CO_EPS0);
    A_0 = CO_HBAR / (CO_ME * CO_C * CO_ALPHA);
    MC2 = CO_ME * (CO_C * CO_C);
    MC = CO_ME * CO_C;
}

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), counter-based replacement for rand()              */
/* ------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define DOM_COLLISION 0u
#define DOM_CHILD_UID 0x5EED0001u
#define DOM_ROULETTE  0x5EED0002u
#define DOM_SPLIT     0x5EED0003u
#define DOM_SHUFFLE   0x5EED0004u

typedef struct {
    uint32_t key[2];
    uint32_t step, seed_lo, seed_hi;
    uint32_t idx;       /* index of the next uniform in this (particle, advance call) stream */
    uint32_t cblock;    /* block currently cached */
    uint32_t cache[4];
    int have;
    const double* inject;   /* replay of reference-emitted vectors: the n-th draw is inject[n] (NULL: Philox) */
    uint32_t ninject;
} rng_t;

static void rng_init(rng_t* g, uint64_t uid, uint32_t domain, uint64_t seed, uint32_t step) {
    g->key[0] = (uint32_t)uid;
    g->key[1] = (uint32_t)(uid >> 32) ^ domain;
    g->step = step;
    g->seed_lo = (uint32_t)seed;
    g->seed_hi = (uint32_t)(seed >> 32);
    g->idx = 0;
    g->have = 0;
    g->cblock = 0;
    g->inject = NULL;
    g->ninject = 0;
}

/* bits -> double in the open interval (0,1): (m + 0.5) * 2^-52, m = top 52 bits of the 64-bit word
 * (exact in binary64; chosen because the device builds it with two shifts and one exact add) */
static inline double bits_to_u01(uint32_t lo, uint32_t hi) {
    uint64_t b = ((uint64_t)hi << 32) | lo;
    return ((double)(b >> 12) + 0.5) * 0x1.0p-52;
}

static double rng_u(rng_t* g) {
    if (g->inject) {        /* the reference's own rand() sequence, recorded by julia/emit_golden.jl */
        double u = g->idx < g->ninject ? g->inject[g->idx] : 0.5;
        g->idx++;
        return u;
    }
    uint32_t block = g->idx >> 1;
    if (!g->have || g->cblock != block) {
        uint32_t ctr[4] = {block, g->step, g->seed_lo, g->seed_hi};
        philox4x32_10(ctr, g->key, g->cache);
        g->cblock = block;
        g->have = 1;
    }
    double u = (g->idx & 1) ? bits_to_u01(g->cache[2], g->cache[3]) : bits_to_u01(g->cache[0], g->cache[1]);
    g->idx++;
    return u;
}

static void child_uids(uint64_t parent, uint32_t idx, uint64_t seed, uint32_t step, uint64_t out[2]) {
    uint32_t key[2] = {(uint32_t)parent, (uint32_t)(parent >> 32) ^ DOM_CHILD_UID};
    uint32_t ctr[4] = {idx, step, (uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    /* uid space: bit 63 set = hashed (births), clear = sequential (assigned by the host side); see ptl_set_uid_counter */
    out[0] = (((uint64_t)o[1] << 32) | o[0]) | PTL_UID_HASHED_BIT;
    out[1] = (((uint64_t)o[3] << 32) | o[2]) | PTL_UID_HASHED_BIT;
}

/* nextcoll() = -log(rand())   src/util.jl:17 */
static inline double nextcoll(rng_t* g) { return -log(rng_u(g)); }

/* ------------------------------------------------------------------------------------ */
/* data structures                                                                       */
/* ------------------------------------------------------------------------------------ */
typedef struct { double v[3]; } vec3;

/* ParticleState: src/electron.jl:13-34, src/photon.jl:6-33, src/positron.jl:3-22 */
typedef struct {
    vec3 x, p;
    double w, t, s, r;
    int active;
} state_t;

typedef struct {
    int ncum, nE;
    double* log_energy;
    double* data; /* [ncum, nE] column-major */
} sb_table;

typedef struct {
    int kind; /* 0 cheb, 1 linear */
    int nprocs;
    ptl_process_desc procs[PTL_MAX_PROCS];
    /* cheb */
    int order, k;
    double xmax;
    double* rate;      /* [order, nprocs, k+1] */
    double* ratebound; /* [order, k+1] */
    /* linear */
    int grid_kind, nE;
    double L1, L2, maxrate;
    double* rbvec;      /* linear tables: vector rate bound on the energy grid (collision_table.jl:35-43) or NULL */
    double* lrate;     /* [nprocs, nE] */
    int64_t counts[PTL_MAX_PROCS + 1];
} table_t;

typedef struct {
    int order, k;
    double xmax;
    double *ec, *pc;
} cheb_loss_t;

typedef struct {
    int used, species, table;
    int64_t capacity, n, iup; /* n, iup: src/population.jl:9-11 (0-based here: iup = first row not yet advanced) */
    double energy_cut;
    double *x, *p, *w, *t, *s, *r; /* x,p: xyz interleaved, like Vector{SVector{3}} */
    uint8_t* active;
    uint64_t* uid;
} pop_t;

typedef struct {
    int npop;
    int pops[PTL_NSPECIES];
    int by_species[PTL_NSPECIES]; /* get(mp, ParticleType) mixed_population.jl:15 */
} multipop_t;

typedef struct {
    int64_t n, cap;
    double *x, *p, *w, *t;
} wallrec_t;

#define MAX_TABLES 16
#define MAX_POPS 64
#define MAX_SB 8
#define MAX_MP 32
#define MAX_CHEBLOSS 4

typedef struct ora_context {
    int ntab, npop, nsb, nmp, ncl;
    table_t tab[MAX_TABLES];
    pop_t pop[MAX_POPS];
    sb_table sb[MAX_SB];
    multipop_t mp[MAX_MP];
    cheb_loss_t cl[MAX_CHEBLOSS];
    uint64_t seed;
    uint32_t step;
    uint64_t next_uid;
    int32_t flags;
    wallrec_t wall[PTL_MAX_WALLS];
    ptl_advance_stats stats;
    char err[256];
} ora_context;

#define FLAG(ctx, bit) do { _Pragma("omp atomic") (ctx)->flags |= (bit); } while (0)

/* ------------------------------------------------------------------------------------ */
/* small vector helpers                                                                  */
/* ------------------------------------------------------------------------------------ */
static inline double dot3(vec3 a, vec3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
static inline vec3 add3(vec3 a, vec3 b) { vec3 r = {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; return r; }
static inline vec3 sub3(vec3 a, vec3 b) { vec3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static inline vec3 scale3(vec3 a, double f) { vec3 r = {{a.v[0] * f, a.v[1] * f, a.v[2] * f}}; return r; }
static inline vec3 div3(vec3 a, double f) { vec3 r = {{a.v[0] / f, a.v[1] / f, a.v[2] / f}}; return r; }
static inline vec3 cross3(vec3 a, vec3 b) {
    vec3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
    return r;
}
/* StaticArrays norm: sqrt(sum of squares) */
static inline double norm3(vec3 a) { return sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]); }
static const vec3 ZERO3 = {{0.0, 0.0, 0.0}};

/* ------------------------------------------------------------------------------------ */
/* kinematics                                                                            */
/* ------------------------------------------------------------------------------------ */
/* kinenergy: electron.jl:53, positron.jl:40, photon.jl:52, slow-electron.jl:27 */
static double kinenergy(int species, vec3 p) {
    switch (species) {
    case PTL_PHOTON: return norm3(p) * CO_C;
    case PTL_SLOW_ELECTRON: return 0.5 * CO_ME * (p.v[0] * p.v[0] + p.v[1] * p.v[1] + p.v[2] * p.v[2]);
    default: return sqrt(MC2 * MC2 + (CO_C * CO_C) * dot3(p, p)) - MC2;
    }
}
/* gamma: electron.jl:54 */
static double gamma_lepton(vec3 p) { return sqrt(1 + (CO_C * CO_C) * dot3(p, p) / (MC2 * MC2)); }
/* velocity: electron.jl:56, positron.jl:43, photon.jl:47 ; slow electron: the `p` column holds v */
static vec3 velocity(int species, vec3 p) {
    switch (species) {
    case PTL_PHOTON: return scale3(p, CO_C / norm3(p));
    case PTL_SLOW_ELECTRON: return p;
    default: return scale3(p, 1 / (CO_ME * gamma_lepton(p)));
    }
}
/* momentum_norm_from_kin: electron.jl:51, positron.jl:38 */
static double pnorm_from_kin(double kin) { return sqrt((kin + MC2) * (kin + MC2) - MC2 * MC2) / CO_C; }
static int charge_of(int species) { return species == PTL_POSITRON ? +1 : (species == PTL_PHOTON ? 0 : -1); }

/* ------------------------------------------------------------------------------------ */
/* turn: src/util.jl:40-57                                                               */
/* ------------------------------------------------------------------------------------ */
static vec3 turn(vec3 u, double cost, double phi, double n) {
    double un = norm3(u);
    vec3 mu = div3(u, un);
    double sinphi = sin(phi), cosphi = cos(phi);
    double st2 = 1 - cost * cost;
    double sint = sqrt(st2 > 0 ? st2 : 0.0); /* guard: reference gives NaN when |cost|>1 by rounding */
    double s2 = 1 - mu.v[2] * mu.v[2];
    double s = sqrt(s2 > 0 ? s2 : 0.0);
    vec3 r;
    if (s == 0.0) { /* guard: reference divides by zero when u is along +-z */
        r.v[0] = sint * cosphi;
        r.v[1] = sint * sinphi;
        r.v[2] = mu.v[2] * cost;
    } else {
        /* b(m1,m2,m3) = m1*m3*cosphi + m2*sinphi */
        double bx = mu.v[0] * mu.v[2] * cosphi + (-mu.v[1]) * sinphi;
        double by = mu.v[1] * mu.v[2] * cosphi + mu.v[0] * sinphi;
        r.v[0] = sint * bx / s + mu.v[0] * cost;
        r.v[1] = sint * by / s + mu.v[1] * cost;
        r.v[2] = -s * sint * cosphi + mu.v[2] * cost;
    }
    return scale3(r, n);
}

/* randsphere: src/util.jl:4-12 */
static vec3 randsphere(rng_t* g) {
    double phi = 2 * CO_PI * rng_u(g);
    double sinphi = sin(phi), cosphi = cos(phi);
    double u = 2 * rng_u(g) - 1;
    double v = sqrt(1 - u * u);
    vec3 r = {{v * cosphi, v * sinphi, u}};
    return r;
}

/* sample_modified_tsai_cos_theta: src/util.jl:143-159 */
static double sample_tsai(rng_t* g, double T) {
    double umax = 2 * (1 + T / (CO_ME * (CO_C * CO_C)));
    double a1 = 1.6, a2 = a1 / 3, border = 0.25;
    double u;
    for (;;) {
        double r1 = rng_u(g), r2 = rng_u(g);
        double uu = -log(r1 * r2);
        u = border > rng_u(g) ? uu * a1 : uu * a2;
        if (u <= umax) break;
    }
    return 1 - 2 * (u * u) / (umax * umax);
}

/* ------------------------------------------------------------------------------------ */
/* Chebyshev lookup: src/cheby.jl:57-81,127-143 ; src/collision_table.jl:82-106          */
/* ------------------------------------------------------------------------------------ */
#define MAX_ORDER 8
typedef struct { int i; double t[MAX_ORDER]; int oob; double w; int kidx; } pre_t;

static void precheb(double x, int k, double xmax, int order, pre_t* pre) {
    double x1 = x / xmax;
    int l;
    double s = frexp(x1, &l);
    int i = (x1 == 0) ? 0 : l + k;
    double xi;
    if (i > 0) {
        xi = 4 * s - 3;
    } else {
        xi = ldexp(1.0, k + 1) * x1 - 1; /* 2^(k+1) * x1 - 1 */
        i = 0;
    }
    pre->oob = 0;
    if (i > k) { i = k; pre->oob = 1; } /* reference indexes out of bounds here (collision_table.jl:91) */
    pre->i = i;
    pre->t[0] = 1.0;
    if (order > 1) pre->t[1] = xi;
    for (int n = 2; n < order; n++) pre->t[n] = 2 * xi * pre->t[n - 1] - pre->t[n - 2];
}

/* sum(ntuple(k -> a[k, ...] * t[k])) — left-to-right, no FMA */
static double chebsum(const double* a, const pre_t* pre, int order) {
    double acc = a[0] * pre->t[0];
    for (int m = 1; m < order; m++) acc = acc + a[m] * pre->t[m];
    return acc;
}

/* ------------------------------------------------------------------------------------ */
/* linear lookup: src/util.jl:23-32,118-127 ; src/collision_table.jl:50-57               */
/* ------------------------------------------------------------------------------------ */
/* Julia LinRange element: r[i] = (1-t)*start + t*stop, t = (i-1)/(len-1)  (base/range.jl lerpi) */
static double linrange_at(double start, double stop, int len, int i1 /*1-based*/) {
    double t = (double)(i1 - 1) / (double)(len - 1);
    return (1 - t) * start + t * stop;
}

static void indweight(const table_t* T, double x, pre_t* pre) {
    double step = (T->L2 - T->L1) / (T->nE - 1); /* step(::LinRange) = (stop-start)/lendiv */
    int i;
    double w;
    if (T->grid_kind == 0) { /* util.jl:23-32 */
        i = (int)floor((x - T->L1) / step) + 1;
        if (i < 1) i = 1;
        if (i > T->nE - 1) { i = T->nE - 1; pre->oob = 1; }
        w = (linrange_at(T->L1, T->L2, T->nE, i + 1) - x) / step;
    } else { /* util.jl:118-127 */
        double x0 = exp(T->L1);
        double l = log(x + x0);
        i = (int)floor((l - T->L1) / step) + 1;
        if (i < 1) i = 1;
        if (i > T->nE - 1) { i = T->nE - 1; pre->oob = 1; }
        double Li1 = linrange_at(T->L1, T->L2, T->nE, i + 1), Li = linrange_at(T->L1, T->L2, T->nE, i);
        w = (exp(Li1) - x0 - x) / (exp(Li1) - exp(Li));
    }
    pre->kidx = i; /* 1-based */
    pre->w = w;
}

static void presample(const table_t* T, double eng, pre_t* pre) {
    pre->oob = 0;
    if (T->kind == 0) precheb(eng, T->k, T->xmax, T->order, pre);
    else indweight(T, eng, pre);
}

static double table_rate(const table_t* T, int j, const pre_t* pre) {
    if (T->kind == 0) {
        const double* a = T->rate + (size_t)T->order * ((size_t)j + (size_t)T->nprocs * pre->i);
        return chebsum(a, pre, T->order);
    }
    int k = pre->kidx - 1;
    return pre->w * T->lrate[j + (size_t)T->nprocs * k] + (1 - pre->w) * T->lrate[j + (size_t)T->nprocs * (k + 1)];
}

static double table_ratebound(ora_context* ctx, const table_t* T, double eng) {
    if (T->kind == 0) {
        pre_t pre;
        precheb(eng, T->k, T->xmax, T->order, &pre);
        if (pre.oob) FLAG(ctx, PTL_ERR_ENERGY_OUT_OF_TABLE);
        return chebsum(T->ratebound + (size_t)T->order * pre.i, &pre, T->order);
    }
    if (T->rbvec) {   /* ratebound(v::Vector, c, eng, pre) collision_table.jl:35-43: the same (k, w) as the rates */
        pre_t pre;
        presample(T, eng, &pre);
        if (pre.oob) FLAG(ctx, PTL_ERR_ENERGY_OUT_OF_TABLE);
        int k = pre.kidx - 1;
        return pre.w * T->rbvec[k] + (1 - pre.w) * T->rbvec[k + 1];
    }
    return T->maxrate; /* ratebound(x::Number, ...) collision_table.jl:33 */
}

/* setr!: src/collisions.jl:63-74 */
static double setr_value(ora_context* ctx, const pop_t* P, vec3 p) {
    double eng = kinenergy(P->species, p);
    if (eng < P->energy_cut) return 0.0;
    return table_ratebound(ctx, &ctx->tab[P->table], eng);
}

/* ------------------------------------------------------------------------------------ */
/* fields and forces: src/field.jl, src/continuum.jl, src/pusher.jl                      */
/* ------------------------------------------------------------------------------------ */
static vec3 eval_field(const ptl_field_desc* f, vec3 x) {
    vec3 r = ZERO3;
    switch (f->kind) {
    case PTL_FIELD_HOMOGENEOUS: /* field.jl:8 */
        r.v[0] = f->par[0]; r.v[1] = f->par[1]; r.v[2] = f->par[2];
        break;
    case PTL_FIELD_DOUBLE_LAYER: /* field.jl:16 */
        if (f->par[0] < x.v[2] && x.v[2] < f->par[1]) { r.v[0] = f->par[2]; r.v[1] = f->par[3]; r.v[2] = f->par[4]; }
        break;
    case PTL_FIELD_STEP: /* field.jl:28 */
        if (x.v[2] < f->par[0]) { r.v[0] = f->par[1]; r.v[1] = f->par[2]; r.v[2] = f->par[3]; }
        else { r.v[0] = f->par[4]; r.v[1] = f->par[5]; r.v[2] = f->par[6]; }
        break;
    case PTL_FIELD_CONFINED_DL: { /* field.jl:42-52 */
        double sx = f->par[0], sy = f->par[1], sz = f->par[2], ez0 = f->par[3];
        double X = x.v[0], Y = x.v[1], Z = x.v[2];
        double ex = exp(-((X * X) / (2 * (sx * sx)) + (Y * Y) / (2 * (sy * sy)) + (Z * Z) / (2 * (sz * sz))));
        r.v[0] = -(ez0 * ex * X * Z) / (sx * sx);
        r.v[1] = -(ez0 * ex * Y * Z) / (sy * sy);
        r.v[2] = (ez0 * ex - (ez0 * ex * (Z * Z)) / (sz * sz));
        break;
    }
    default: break;
    }
    return r;
}

/* x0x1: continuum.jl:123-139 */
static void x0x1(double C, double* x0, double* x1) {
    if (C < 10) { *x0 = 1.6; *x1 = 4.0; }
    else if (C < 10.5) { *x0 = 1.7; *x1 = 4.0; }
    else if (C < 11.0) { *x0 = 1.8; *x1 = 4.0; }
    else if (C < 11.5) { *x0 = 1.9; *x1 = 4.0; }
    else if (C < 12.25) { *x0 = 2.0; *x1 = 4.0; }
    else if (C < 13.804) { *x0 = 2.0; *x1 = 5.0; }
    else { *x0 = 0.326 * C - 2.5; *x1 = 5.0; }
}

/* energy_loss: continuum.jl:63-96 ; _F: :107-120 ; taumax: :102-103 */
static double energy_loss(double nel, double I, double Tcut, int species, double eng) {
    double tau = eng / MC2, tauc = Tcut / MC2;
    double taumax = (species == PTL_POSITRON) ? tau : tau / 2;
    double gam = 1 + tau;
    double beta2 = 1 - 1 / (gam * gam);
    double tauup = tauc < taumax ? tauc : taumax;
    double F;
    if (species == PTL_POSITRON) {
        double y = 1 / (2 + tau);
        double tu = tauup;
        F = (log(tau * tu) - ((tu * tu) / tau) * (tau * 2 * tu - 3 * (tu * tu) * y / 2 - (tu - (tu * tu * tu) / 3) * (y * y)
                                                   - ((tu * tu) / 2 - tau * (tu * tu * tu) / 3 + (tu * tu * tu * tu) / 4) * (y * y * y)));
    } else {
        double tu = tauup;
        F = (-1 - beta2 + log((tau - tu) * tu) + tau / (tau - tu) + ((tu * tu) / 2 + (2 * tau + 1) * log(1 - tu / tau)) / (gam * gam));
    }
    double x = log((gam * gam) * beta2) / log(10.0) / 2;
    double hnup = CO_HBAR * CO_C * sqrt(4 * CO_PI * nel * R_E);
    double C = 1 + 2 * log(I / hnup);
    double xa = C / log(10.0) / 2;
    double x0, x1;
    x0x1(C, &x0, &x1);
    double d = (x1 - x0);
    double a = 2 * log(10.0) * (xa - x) / (d * d * d);
    double delta;
    if (x < x0) delta = 0.0;
    else if (x < x1) { double e = (x1 - x); delta = 2 * log(10.0) * x - C + a * (e * e * e); }
    else delta = 2 * log(10.0) * x - C;
    double IM = I / MC2;
    return (2 * CO_PI * (R_E * R_E) * MC2 * nel / beta2) * (log((2 * (gam + 1)) / (IM * IM)) + F - delta);
}

static int mask_has(uint32_t mask, int species) { return mask == 0 || ((mask >> species) & 1u); }

/* force(forcing, s): pusher.jl:8-34 ; field.jl:62-70 ; continuum.jl:17-22,45-57 */
static vec3 total_force(const ora_context* ctx, const ptl_pusher_desc* psh, int species, vec3 x, vec3 p) {
    vec3 acc = ZERO3;
    /* _force(tpl) = force(first) + _force(tail): right-nested sum, terminating in zero */
    for (int k = psh->nforcings - 1; k >= 0; k--) {
        const ptl_forcing_desc* f = &psh->forcing[k];
        vec3 fk = ZERO3;
        if (mask_has(f->species_mask, species)) {
            switch (f->kind) {
            case PTL_FORCE_EM:
                if (species != PTL_PHOTON) {
                    vec3 e = eval_field(&f->e, x);
                    vec3 b = eval_field(&f->b, x);
                    vec3 v = velocity(species, p);
                    double q = charge_of(species) * CO_E;
                    fk = scale3(add3(e, cross3(v, b)), q);
                    if (species == PTL_SLOW_ELECTRON) fk = div3(fk, CO_ME); /* the `p` column holds v: dv/dt = F/m */
                }
                break;
            case PTL_FORCE_CONTINUUM:
                if (species == PTL_ELECTRON || species == PTL_POSITRON) {
                    double fl = energy_loss(f->nel, f->I, f->Tcut, species, kinenergy(species, p));
                    fk = scale3(p, -fl / norm3(p));
                }
                break;
            case PTL_FORCE_CHEB_CONTINUUM:
                if (species == PTL_ELECTRON || species == PTL_POSITRON) {
                    const cheb_loss_t* cl = &ctx->cl[f->cheb_id];
                    pre_t pre;
                    precheb(kinenergy(species, p), cl->k, cl->xmax, cl->order, &pre);
                    const double* a = (species == PTL_ELECTRON ? cl->ec : cl->pc) + (size_t)cl->order * pre.i;
                    double fl = chebsum(a, &pre, cl->order);
                    fk = scale3(p, -fl / norm3(p));
                }
                break;
            default: break;
            }
        }
        acc = add3(fk, acc);
    }
    return acc;
}

/* advance_particle: pusher.jl:41-63 (RK2Pusher), :67-73 (RestrictedPusher), :75-76 (NullPusher) */
static state_t advance_particle(const ora_context* ctx, const ptl_pusher_desc* psh, int species, state_t y, double dt) {
    state_t yf = y;
    if (psh->kind == PTL_PUSHER_NULL || !mask_has(psh->restrict_mask, species)) {
        yf.t = y.t + dt;
        return yf;
    }
    vec3 v1 = velocity(species, y.p);
    vec3 f1 = total_force(ctx, psh, species, y.x, y.p);
    vec3 x2 = add3(y.x, div3(scale3(v1, 2 * dt), 3));
    vec3 p2 = add3(y.p, div3(scale3(f1, 2 * dt), 3));
    vec3 v2 = velocity(species, p2);
    vec3 f2 = total_force(ctx, psh, species, x2, p2);
    yf.x = add3(y.x, scale3(add3(div3(v1, 4), div3(scale3(v2, 3), 4)), dt));
    yf.p = add3(y.p, scale3(add3(div3(f1, 4), div3(scale3(f2, 3), 4)), dt));
    yf.t = y.t + dt;
    return yf;
}

/* ------------------------------------------------------------------------------------ */
/* collision outcomes: src/collisions.jl:11-55                                           */
/* ------------------------------------------------------------------------------------ */
enum { OUT_NULL = 0, OUT_STATE_CHANGE, OUT_NEW_PARTICLE, OUT_REMOVE, OUT_REPLACE, OUT_REPLACE_PAIR };

typedef struct {
    int kind;
    state_t s1;          /* new state of the colliding particle (STATE_CHANGE, NEW_PARTICLE) */
    int sp2, sp3;        /* species of state2 / state3 */
    state_t s2, s3;
} outcome_t;

/* 4-arg state constructor: w, t given; s = nextcoll() drawn HERE; r = 0; active = true
 * (electron.jl:30-33, photon.jl:26-31, positron.jl:18-21) */
static state_t new_state(rng_t* g, vec3 x, vec3 p, double w, double t) {
    state_t s;
    s.x = x; s.p = p; s.w = w; s.t = t;
    s.s = nextcoll(g);
    s.r = 0.0;
    s.active = 1;
    return s;
}

/* Lehtinen 1999 two-body angles shared by RBEB / Moller / Bhaba (rbeb.jl:65-80) */
static void ionization_products(rng_t* g, const state_t* st, double E0, double E1, double E2, int sp_primary, outcome_t* out) {
    double p1 = sqrt(E1 * E1 + 2 * MC2 * E1) / CO_C;
    double p2 = sqrt(E2 * E2 + 2 * MC2 * E2) / CO_C;
    double cos1 = sqrt(E1 * (E0 + 2 * MC2) / (E0 * (E1 + 2 * MC2)));
    double cos2 = sqrt(E2 * (E0 + 2 * MC2) / (E0 * (E2 + 2 * MC2)));
    double phi = 2 * CO_PI * rng_u(g);
    vec3 p1v = turn(st->p, cos1, phi, p1);
    vec3 p2v = turn(st->p, cos2, -phi, p2);
    out->kind = OUT_NEW_PARTICLE;
    out->s1 = new_state(g, st->x, p1v, st->w, st->t);
    out->sp2 = PTL_ELECTRON;
    out->s2 = new_state(g, st->x, p2v, st->w, st->t);
    (void)sp_primary;
}

/* RelativisticCoulomb: relativistic_coulomb.jl:10-51 */
static void collide_coulomb(rng_t* g, const ptl_process_desc* pr, int species, const state_t* st, outcome_t* out) {
    double phi = 2 * CO_PI * rng_u(g);
    double beta = norm3(velocity(species, st->p)) / CO_C;
    double a = 1.3413 * pow(pr->par[0], -1.0 / 3.0) * A_0;
    double alpha = (CO_HBAR * CO_HBAR) / (4 * dot3(st->p, st->p) * (a * a));
    double x;
    for (;;) { /* sample_rel_sr :35-51 */
        double u = rng_u(g);
        x = alpha * u / (alpha + 1 - u);
        double z = rng_u(g);
        if (z < (1 - (beta * beta) * x)) break;
    }
    double cost = 1 - 2 * x;
    vec3 pnew = turn(st->p, cost, phi, norm3(st->p));
    out->kind = OUT_STATE_CHANGE;
    out->s1 = new_state(g, st->x, pnew, st->w, st->t);
}

/* RBEB: rbeb.jl:54-80 (collide), :156-196 (sampler) */
static void collide_rbeb(ora_context* ctx, rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    double B = pr->par[0];
    double mc2 = CO_ME * (CO_C * CO_C);
    double t1 = eng / mc2, b1 = B / mc2;
    double bt2 = 1 - 1 / ((1 + t1) * (1 + t1));
    double t = eng / B;
    double A = -(1 + 2 * t1) / (t + 1) / ((1 + t1) * (1 + t1));
    double Bc = 1;
    double C = (log(bt2 / (1 - bt2)) - bt2 - log(2 * b1));
    double M = (b1 * b1) / ((1 + t1) * (1 + t1));
    double w;
    for (;;) {
        double u = rng_u(g);
        w = u / ((t + 1) / (t - 1) - u);
        double pb = (2 * Bc + 2 * C + (t + 1) * (t + 1) * Bc * M / 4) / ((1 + w) * (1 + w));
        /* g(t,w,q) = 1/(w+1)^q + 1/(t-w)^q */
        double g1 = 1 / (w + 1) + 1 / (t - w);
        double g2 = 1 / ((w + 1) * (w + 1)) + 1 / ((t - w) * (t - w));
        double g3 = 1 / ((w + 1) * (w + 1) * (w + 1)) + 1 / ((t - w) * (t - w) * (t - w));
        double p0 = A * g1 + Bc * (g2 + M) + C * g3;
        if (rng_u(g) * pb < p0) break;
    }
    double E2 = B * w;
    double E0 = eng;
    double E1 = E0 - E2 - B;
    if (!(E2 < E1)) FLAG(ctx, PTL_ERR_SAMPLER_INVARIANT); /* @assert E2 < E1  rbeb.jl:63 */
    ionization_products(g, st, E0, E1, E2, PTL_ELECTRON, out);
}

/* Moller: moller.jl:13-37 (collide), :64-87 (sampler) */
static void collide_moller(rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    double tcut = pr->par[1];
    double eps0 = tcut / eng;
    double gam = 1 + eng / MC2;
    double eps;
    for (;;) {
        double r = rng_u(g);
        eps = eps0 / (1 - r + 2 * eps0 * r);
        double gg = 4 / (9 * (gam * gam) - 10 * gam + 5) *
                    ((gam - 1) * (gam - 1) * (eps * eps) - (2 * (gam * gam) + 2 * gam - 1) * (eps / (1 - eps)) + (gam * gam) / ((1 - eps) * (1 - eps)));
        if (rng_u(g) < gg) break;
    }
    double E2 = eps * eng;
    ionization_products(g, st, eng, eng - E2, E2, PTL_ELECTRON, out);
}

/* bhaba_bs: bhaba.jl:84-91 */
static void bhaba_bs(double y, double B[5]) {
    double q = (1 - 2 * y);
    B[1] = 2 - y * y;
    B[2] = q * (3 + y * y);
    B[3] = q * q + q * q * q;
    B[4] = q * q * q;
}

/* Bhaba: bhaba.jl:9-33 (collide), :55-81 (sampler) */
static void collide_bhaba(rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    double tcut = pr->par[1];
    double eps0 = tcut / eng;
    double gam = 1 + eng / MC2;
    double y = 1 / (gam + 1);
    double B[5];
    B[0] = (gam * gam) / ((gam * gam) - 1);
    bhaba_bs(y, B);
    double eps;
    for (;;) {
        double r = rng_u(g);
        eps = eps0 / (1 - r + eps0 * r);
        double g1 = B[0] + B[1] * eps + B[2] * (eps * eps) + B[3] * (eps * eps * eps) + B[4] * (eps * eps * eps * eps);
        double g2 = B[0] + B[1] * eps0 + B[2] * (eps0 * eps0) + B[3] * (eps0 * eps0 * eps0) + B[4] * (eps0 * eps0 * eps0 * eps0);
        if (rng_u(g) < g1 / g2) break;
    }
    double E2 = eps * eng;
    ionization_products(g, st, eng, eng - E2, E2, PTL_POSITRON, out);
}

/* first index (1-based) with a[i] >= x in an ascending vector; n+1 if none (Base.searchsortedfirst) */
static int searchsortedfirst(const double* a, int n, double x) {
    int lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        int m = lo + ((hi - lo) >> 1);
        if (a[m - 1] < x) lo = m; else hi = m;
    }
    return hi;
}

/* SeltzerBerger: seltzer.jl:67-90 (collide), :97-122 (sampler) */
static void collide_seltzer(ora_context* ctx, rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    const sb_table* sb = &ctx->sb[pr->aux];
    double x = rng_u(g);
    double y = log(eng);
    int nc = sb->ncum;
    /* pcum = LinRange(0,1,ncum): pcum[i] = (i-1)/(ncum-1).  searchsortedfirst = first i with pcum[i] >= x */
    int i2 = (int)ceil(x * (nc - 1)) + 1;
    if (i2 < 2) i2 = 2;
    if (i2 > nc) i2 = nc;
    while (i2 > 2 && linrange_at(0.0, 1.0, nc, i2 - 1) >= x) i2--;
    while (i2 < nc && linrange_at(0.0, 1.0, nc, i2) < x) i2++;
    int i1 = i2 - 1;
    int j2 = searchsortedfirst(sb->log_energy, sb->nE, y);
    if (j2 < 2) { j2 = 2; FLAG(ctx, PTL_ERR_ENERGY_OUT_OF_TABLE); }
    if (j2 > sb->nE) { j2 = sb->nE; FLAG(ctx, PTL_ERR_ENERGY_OUT_OF_TABLE); }
    int j1 = j2 - 1;
    double x1 = linrange_at(0.0, 1.0, nc, i1), x2 = linrange_at(0.0, 1.0, nc, i2);
    double y1 = sb->log_energy[j1 - 1], y2 = sb->log_energy[j2 - 1];
    const double* u = sb->data;
#define U(i, j) u[((i)-1) + (size_t)nc * ((j)-1)]
    double A = (x2 - x1) * (y2 - y1);
    double S = (U(i1, j1) * (x2 - x) * (y2 - y) + U(i1, j2) * (x - x1) * (y2 - y) + U(i2, j1) * (x2 - x) * (y - y1) + U(i2, j2) * (x - x1) * (y - y1));
#undef U
    double k = eng * exp(S / A);
    if (!(k < eng)) FLAG(ctx, PTL_ERR_SAMPLER_INVARIANT); /* seltzer.jl:73 */
    double pph = k / CO_C;
    double cost = sample_tsai(g, eng);
    double phi = 2 * CO_PI * rng_u(g);
    vec3 p_ph = turn(st->p, cost, phi, pph);
    vec3 p_e = sub3(st->p, p_ph);
    out->kind = OUT_NEW_PARTICLE;
    out->s1 = new_state(g, st->x, p_e, st->w, st->t);
    out->sp2 = PTL_PHOTON;
    out->s2 = new_state(g, st->x, p_ph, st->w, st->t);
}

/* Compton: compton.jl:9-28 (collide), :119-144 (sampler) */
static void collide_compton(rng_t* g, const state_t* st, double eng, outcome_t* out) {
    double eps0 = MC2 / (MC2 + 2 * eng);
    double a1 = -log(eps0);
    double a2 = (1 - eps0 * eps0) / 2;
    double t, eps;
    for (;;) {
        if (rng_u(g) < a1 / (a1 + a2)) eps = exp(-rng_u(g) * a1);
        else eps = sqrt(eps0 * eps0 + (1 - eps0 * eps0) * rng_u(g));
        t = MC2 * (1 - eps) / (eps * eng);
        double gg = (1 - eps / (1 + eps * eps) * t * (2 - t));
        if (rng_u(g) < gg) break;
    }
    double cost = 1 - t;
    double E1 = eps * eng;
    double phi = 2 * CO_PI * rng_u(g);
    vec3 pg = turn(st->p, cost, phi, E1 / CO_C);
    vec3 pe = sub3(st->p, pg);
    out->kind = OUT_NEW_PARTICLE;
    out->s1 = new_state(g, st->x, pg, st->w, st->t);
    out->sp2 = PTL_ELECTRON;
    out->s2 = new_state(g, st->x, pe, st->w, st->t);
}

/* PhotoElectric: photo_electric.jl:38-52 (collide), :60-76 (energy), :78-99 (angle) */
static void collide_photoelectric(ora_context* ctx, rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    int nb = (int)pr->par[1];
    double b = 0;
    for (int i = 0; i < nb; i++) {
        b = pr->par[2 + i];
        if (eng > b) break;
    }
    if (!(eng > b)) FLAG(ctx, PTL_ERR_SAMPLER_INVARIANT); /* photo_electric.jl:71 */
    double Ee = eng - b;
    double gam = 1 + Ee / MC2;
    double beta = sqrt(1 - 1 / (gam * gam));
    double A = 1 / beta - 1;
    double K = beta * gam * (gam - 1) * (gam - 2) / 2;
    double g0 = (2 - 0.0) * (1 / (A + 0.0) + K);
    double nu;
    for (;;) {
        double xi = rng_u(g);
        nu = 2 * A / ((A + 2) * (A + 2) - 4 * xi) * (2 * xi + (A + 2) * sqrt(xi));
        double xi1 = rng_u(g);
        double gn = (2 - nu) * (1 / (A + nu) + K);
        if (xi1 * g0 < gn) break;
    }
    double cost = 1 - nu;
    double phi = 2 * CO_PI * rng_u(g);
    double pn = pnorm_from_kin(Ee);
    vec3 p = turn(st->p, cost, phi, pn);
    out->kind = OUT_REPLACE;
    out->sp2 = PTL_ELECTRON;
    out->s2 = new_state(g, st->x, p, st->w, st->t);
}

/* screen functions: bethe_heitler.jl:163-193 */
static double bh_screen1(double d) { return d > 1.4 ? 42.038 - 8.29 * log(d + 0.958) : 42.184 - d * (7.444 - 1.623 * d); }
static double bh_screen2(double d) { return d > 1.4 ? 42.038 - 8.29 * log(d + 0.958) : 41.326 - d * (5.848 - 0.902 * d); }
/* _fc: bethe_heitler.jl:153-160 (alphaZ = fine_structure, Z not multiplied: replicated) */
static double bh_fc(void) {
    double aZ = CO_ALPHA, aZ2 = aZ * aZ, aZ4 = aZ2 * aZ2, aZ6 = aZ4 * aZ2;
    double f1 = 1 / (1 + aZ2) + 0.20206 - 0.0369 * aZ2 + 0.0083 * aZ4 - 0.0020 * aZ6;
    return f1 * aZ2;
}

/* BetheHeitler: bethe_heitler.jl:5-25 (collide), :85-146 (sampler) */
static void collide_bethe_heitler(ora_context* ctx, rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    double Z = pr->par[0];
    double eps0 = MC2 / eng;
    if (!(eps0 < 0.5)) FLAG(ctx, PTL_ERR_SAMPLER_INVARIANT); /* bethe_heitler.jl:90 */
    double eps;
    if (eng < 2e6 * CO_E) {
        eps = eps0 + (0.5 - eps0) * rng_u(g);
    } else {
        double d0 = 136 * eps0 / pow(Z, 1.0 / 3.0);
        double FZ = 8 * log(Z) / 3;
        if (eng > 50e6 * CO_E) FZ += 8 * bh_fc();
        double dmin = 4 * d0;
        double dmax = exp((42.24 - FZ) / 8.368) - 0.952;
        double epsp = (1 - sqrt(1 - dmin / dmax)) / 2;
        double epsmin = eps0 > epsp ? eps0 : epsp;
        double epsrange = 0.5 - epsmin;
        double F10 = bh_screen1(dmin), F20 = bh_screen2(dmin);
        F10 -= FZ;
        F20 -= FZ;
        double NF1 = F10 * (epsrange * epsrange); if (!(NF1 > 0)) NF1 = 0;
        double NF2 = 1.5 * F20; if (!(NF2 > 0)) NF2 = 0;
        double NC = NF1 / (NF1 + NF2);
        for (;;) {
            if (NC > rng_u(g)) {
                eps = 0.5 - epsrange * pow(rng_u(g), 1.0 / 3.0);
                double d = d0 / (eps * (1 - eps));
                if (rng_u(g) < (bh_screen1(d) - FZ) / F10) break;
            } else {
                eps = epsmin + epsrange * rng_u(g);
                double d = d0 / (eps * (1 - eps));
                if (rng_u(g) < (bh_screen2(d) - FZ) / F20) break;
            }
        }
    }
    double etot, ptot;
    if (rng_u(g) < 0.5) { etot = (1 - eps) * eng; ptot = eps * eng; } /* rand(Bool) */
    else { ptot = (1 - eps) * eng; etot = eps * eng; }
    double ekin_ret = etot - MC2 > 0 ? etot - MC2 : 0.0;
    double pkin_ret = ptot - MC2 > 0 ? ptot - MC2 : 0.0;
    /* collide destructures `pkin, ekin = sample_secondary_energy(...)` which returns (ekin, pkin)
     * (bethe_heitler.jl:6 vs :145) — replicated literally */
    double pkin = ekin_ret, ekin = pkin_ret;
    double phi = 2 * CO_PI * rng_u(g);
    double cost = sample_tsai(g, ekin);
    vec3 p_e = turn(st->p, cost, phi, pnorm_from_kin(ekin));
    cost = sample_tsai(g, pkin);
    vec3 p_p = turn(st->p, cost, phi, pnorm_from_kin(pkin));
    out->kind = OUT_REPLACE_PAIR;
    out->sp2 = PTL_ELECTRON;
    out->s2 = new_state(g, st->x, p_e, st->w, st->t);
    out->sp3 = PTL_POSITRON;
    out->s3 = new_state(g, st->x, p_p, st->w, st->t);
}

/* PositronAnihilation: anihilation.jl:6-23 (collide), :39-67 (sampler, angle) */
static void collide_anihilation(rng_t* g, const state_t* st, double eng, outcome_t* out) {
    double gam = 1 + eng / MC2;
    double sq = sqrt((gam - 1) / (gam + 1));
    double epsmax = (1 + sq) / 2, epsmin = (1 - sq) / 2;
    double eps;
    for (;;) {
        eps = epsmin * pow(epsmax / epsmin, rng_u(g));
        double gg = 1 - eps + (2 * gam * eps - 1) / (eps * ((gam + 1) * (gam + 1)));
        if (rng_u(g) < gg) break;
    }
    double cost = (eps * (gam + 1) - 1) / (eps * sqrt(gam * gam - 1));
    double phi = 2 * CO_PI * rng_u(g);
    double Etot = eng + 2 * MC2;
    double pan = eps * Etot / CO_C;
    vec3 pa = turn(st->p, cost, phi, pan);
    vec3 pb = sub3(st->p, pa);
    out->kind = OUT_REPLACE_PAIR;
    out->sp2 = PTL_PHOTON;
    out->s2 = new_state(g, st->x, pa, st->w, st->t);
    out->sp3 = PTL_PHOTON;
    out->s3 = new_state(g, st->x, pb, st->w, st->t);
}

/* LXCat kinds: slow-electron.jl:108-139, read consistently with the current loop: outgoing
 * states get a fresh s (one draw each, after the direction draws); the `p` column holds v. */
static void collide_lx(rng_t* g, const ptl_process_desc* pr, const state_t* st, double eng, outcome_t* out) {
    switch (pr->kind) {
    case PTL_PROC_LX_EXCITATION: {
        double E1 = eng - pr->par[0]; if (!(E1 > 0)) E1 = 0;
        double vabs = sqrt(2 * E1 / CO_ME);
        vec3 v1 = scale3(randsphere(g), vabs);
        out->kind = OUT_STATE_CHANGE;
        out->s1 = new_state(g, st->x, v1, st->w, st->t);
        break;
    }
    case PTL_PROC_LX_IONIZATION: {
        double E1 = eng - pr->par[0]; if (!(E1 > 0)) E1 = 0;
        E1 = E1 / 2;
        double vabs = sqrt(2 * E1 / CO_ME);
        vec3 v = scale3(randsphere(g), vabs);
        vec3 v1 = scale3(randsphere(g), vabs);
        out->kind = OUT_NEW_PARTICLE;
        out->s1 = new_state(g, st->x, v, st->w, st->t);
        out->sp2 = PTL_SLOW_ELECTRON;
        out->s2 = new_state(g, st->x, v1, st->w, st->t);
        break;
    }
    case PTL_PROC_LX_ATTACHMENT:
        out->kind = OUT_REMOVE;
        break;
    case PTL_PROC_LX_ELASTIC: {
        double mr = pr->par[0];
        vec3 vcm = scale3(st->p, mr / (1 + mr));
        vec3 d = sub3(st->p, vcm);
        vec3 vf = add3(scale3(randsphere(g), norm3(d)), vcm);
        out->kind = OUT_STATE_CHANGE;
        out->s1 = new_state(g, st->x, vf, st->w, st->t);
        break;
    }
    default: out->kind = OUT_NULL; break;
    }
}

static void collide(ora_context* ctx, rng_t* g, const ptl_process_desc* pr, int species, const state_t* st, double eng, outcome_t* out) {
    switch (pr->kind) {
    case PTL_PROC_COULOMB: collide_coulomb(g, pr, species, st, out); break;
    case PTL_PROC_RBEB: collide_rbeb(ctx, g, pr, st, eng, out); break;
    case PTL_PROC_MOLLER: collide_moller(g, pr, st, eng, out); break;
    case PTL_PROC_BHABA: collide_bhaba(g, pr, st, eng, out); break;
    case PTL_PROC_SELTZER: collide_seltzer(ctx, g, pr, st, eng, out); break;
    case PTL_PROC_COMPTON: collide_compton(g, st, eng, out); break;
    case PTL_PROC_PHOTOELECTRIC: collide_photoelectric(ctx, g, pr, st, eng, out); break;
    case PTL_PROC_BETHE_HEITLER: collide_bethe_heitler(ctx, g, pr, st, eng, out); break;
    case PTL_PROC_ANIHILATION: collide_anihilation(g, st, eng, out); break;
    case PTL_PROC_LX_EXCITATION:
    case PTL_PROC_LX_IONIZATION:
    case PTL_PROC_LX_ATTACHMENT:
    case PTL_PROC_LX_ELASTIC: collide_lx(g, pr, st, eng, out); break;
    default: out->kind = OUT_NULL; break; /* collide(::NullCollision) collisions.jl:58 */
    }
}

/* ------------------------------------------------------------------------------------ */
/* the store: src/population.jl                                                          */
/* ------------------------------------------------------------------------------------ */
static state_t load_state(const pop_t* P, int64_t i) {
    state_t s;
    for (int c = 0; c < 3; c++) { s.x.v[c] = P->x[3 * i + c]; s.p.v[c] = P->p[3 * i + c]; }
    s.w = P->w[i]; s.t = P->t[i]; s.s = P->s[i]; s.r = P->r[i]; s.active = P->active[i];
    return s;
}
static void store_state(pop_t* P, int64_t i, const state_t* s) {
    for (int c = 0; c < 3; c++) { P->x[3 * i + c] = s->x.v[c]; P->p[3 * i + c] = s->p.v[c]; }
    P->w[i] = s->w; P->t[i] = s->t; P->s[i] = s->s; P->r[i] = s->r; P->active[i] = (uint8_t)(s->active != 0);
}

/* add_particle!: population.jl:103-113.  Returns new row (0-based) or -1. */
static int64_t add_particle(ora_context* ctx, pop_t* P, const state_t* s, uint64_t uid) {
    if (kinenergy(P->species, s->p) <= P->energy_cut) return -1;
    int64_t slot;
#pragma omp atomic capture
    slot = P->n++;
    if (slot >= P->capacity) {
#pragma omp atomic
        P->n--;
        FLAG(ctx, PTL_ERR_CAPACITY_OVERFLOW);
        return -1;
    }
    store_state(P, slot, s);
    P->uid[slot] = uid;
    return slot;
}

/* apply!: collisions.jl:83-133 */
static void apply_outcome(ora_context* ctx, const multipop_t* mp, pop_t* P, int64_t i, const outcome_t* o, rng_t* g,
                          int64_t* births) {
    uint64_t cu[2];
    switch (o->kind) {
    case OUT_NULL: /* :83-88 */
        P->r[i] = setr_value(ctx, P, load_state(P, i).p);
        P->s[i] = nextcoll(g);
        break;
    case OUT_STATE_CHANGE: /* :90-94 */
        store_state(P, i, &o->s1);
        P->r[i] = setr_value(ctx, P, o->s1.p);
        break;
    case OUT_NEW_PARTICLE: { /* :96-105 */
        store_state(P, i, &o->s1);
        P->r[i] = setr_value(ctx, P, o->s1.p);
        child_uids(P->uid[i], g->idx, ctx->seed, ctx->step, cu);
        if (mp->by_species[o->sp2] < 0) break;
        pop_t* P2 = &ctx->pop[mp->by_species[o->sp2]];
        int64_t j = add_particle(ctx, P2, &o->s2, cu[0]);
        if (j >= 0) { P2->r[j] = setr_value(ctx, P2, o->s2.p); (*births)++; }
        break;
    }
    case OUT_REMOVE: /* :107-110 */
        P->active[i] = 0;
        break;
    case OUT_REPLACE: { /* :112-120 */
        P->active[i] = 0;
        child_uids(P->uid[i], g->idx, ctx->seed, ctx->step, cu);
        if (mp->by_species[o->sp2] < 0) break;
        pop_t* P2 = &ctx->pop[mp->by_species[o->sp2]];
        int64_t j = add_particle(ctx, P2, &o->s2, cu[0]);
        if (j >= 0) { P2->r[j] = setr_value(ctx, P2, o->s2.p); (*births)++; }
        break;
    }
    case OUT_REPLACE_PAIR: { /* :122-133 */
        P->active[i] = 0;
        child_uids(P->uid[i], g->idx, ctx->seed, ctx->step, cu);
        if (mp->by_species[o->sp2] >= 0) {
            pop_t* P2 = &ctx->pop[mp->by_species[o->sp2]];
            int64_t j = add_particle(ctx, P2, &o->s2, cu[0]);
            if (j >= 0) { P2->r[j] = setr_value(ctx, P2, o->s2.p); (*births)++; }
        }
        if (mp->by_species[o->sp3] >= 0) {
            pop_t* P3 = &ctx->pop[mp->by_species[o->sp3]];
            int64_t j = add_particle(ctx, P3, &o->s3, cu[1]);
            if (j >= 0) { P3->r[j] = setr_value(ctx, P3, o->s3.p); (*births)++; }
        }
        break;
    }
    }
}

/* do_one_collision!: collisions.jl:142-199 */
static void do_one_collision(ora_context* ctx, const multipop_t* mp, pop_t* P, const state_t* st, int64_t i, rng_t* g,
                             int count, int64_t* births) {
    table_t* T = &ctx->tab[P->table];
    if (st->r == 0) return;                 /* :148 */
    double eng = kinenergy(P->species, st->p);
    if (!(eng >= P->energy_cut)) return;    /* :151 */
    pre_t pre;
    presample(T, eng, &pre);                /* :153 */
    if (pre.oob) FLAG(ctx, PTL_ERR_ENERGY_OUT_OF_TABLE);
    /* stream layout: every collision test starts on an even draw index, so that the two halves of a Philox
     * block are consumed pairwise in the same order by every particle (keeps the device lanes convergent) */
    g->idx += g->idx & 1u;
    double xi = rng_u(g) * st->r;           /* :154 */
    outcome_t out;
    for (int j = 0; j < T->nprocs; j++) {   /* :166-180, unrolled in the reference */
        double nu = table_rate(T, j, &pre);
        if (nu > xi) {
            collide(ctx, g, &T->procs[j], P->species, st, eng, &out);
            if (count) {
#pragma omp atomic
                T->counts[j]++;
            }
            apply_outcome(ctx, mp, P, i, &out, g, births);
            return;
        } else {
            xi -= nu;
        }
    }
    if (!(xi >= 0)) FLAG(ctx, PTL_ERR_RATE_BOUND_VIOLATED); /* :186 */
    if (count) {
#pragma omp atomic
        T->counts[T->nprocs]++;
    }
    out.kind = OUT_NULL;
    apply_outcome(ctx, mp, P, i, &out, g, births);
}

/* WallCallback.onadvance: callback.jl:167-184 ; lincomb: electron.jl:127-132 */
static void wall_push(ora_context* ctx, int iw, vec3 x, vec3 p, double w, double t) {
    wallrec_t* W = &ctx->wall[iw];
#pragma omp critical(wallrec)
    {
        if (W->n == W->cap) {
            int64_t nc = W->cap ? 2 * W->cap : 1024;
            W->x = realloc(W->x, sizeof(double) * 3 * nc);
            W->p = realloc(W->p, sizeof(double) * 3 * nc);
            W->w = realloc(W->w, sizeof(double) * nc);
            W->t = realloc(W->t, sizeof(double) * nc);
            W->cap = nc;
        }
        int64_t k = W->n++;
        for (int c = 0; c < 3; c++) { W->x[3 * k + c] = x.v[c]; W->p[3 * k + c] = p.v[c]; }
        W->w[k] = w; W->t[k] = t;
    }
}

static state_t onadvance(ora_context* ctx, const ptl_callback_desc* cb, int species, const state_t* old, state_t nw, rng_t* g) {
    if (!cb) return nw;
    for (int k = 0; k < cb->nwalls; k++) { /* CombinedCallback: callback.jl:71-81 */
        const ptl_wall_desc* wd = &cb->wall[k];
        if (wd->species != species) continue;
        int c = wd->coord;
        if (old->x.v[c] < wd->v && wd->v < nw.x.v[c]) {
            double w = (wd->v - old->x.v[c]) / (nw.x.v[c] - old->x.v[c]);
            /* mid = lincomb(new, old, w): a*w + b*(1-w), 4-arg ctor => one discarded s draw */
            vec3 mx = add3(scale3(nw.x, w), scale3(old->x, 1 - w));
            vec3 mp = add3(scale3(nw.p, w), scale3(old->p, 1 - w));
            double mw = nw.w * w + old->w * (1 - w);
            double mt = nw.t * w + old->t * (1 - w);
            (void)rng_u(g);
            wall_push(ctx, k, mx, mp, mw, mt);
            if (wd->drop) nw.active = 0;
        }
    }
    return nw;
}

/* advance1!: mixed_population.jl:56-93 */
static int64_t advance1(ora_context* ctx, const multipop_t* mp, const ptl_pusher_desc* psh, double tfinal,
                        const ptl_callback_desc* cb) {
    int64_t total = 0;
    int count = cb ? cb->count_collisions : 0;
    for (int ip = 0; ip < mp->npop; ip++) {
        pop_t* P = &ctx->pop[mp->pops[ip]];
        int64_t ilast = P->n; /* rows [iup, ilast) */
        int64_t iup = P->iup;
        int64_t substeps = 0, births = 0;
#pragma omp parallel for schedule(static) reduction(+ : substeps, births)
        for (int64_t i = iup; i < ilast; i++) {
            if (!P->active[i]) continue;
            rng_t g;
            rng_init(&g, P->uid[i], DOM_COLLISION, ctx->seed, ctx->step);
            double trem = tfinal - P->t[i];
            while (trem > DBL_EPSILON && P->active[i]) {
                double tnext = P->s[i] / P->r[i];
                double dt;
                int collides;
                if (trem > tnext) { dt = tnext; collides = 1; }
                else { dt = trem; collides = 0; P->s[i] -= dt * P->r[i]; }
                state_t st = load_state(P, i);
                state_t nw = advance_particle(ctx, psh, P->species, st, dt);
                nw = onadvance(ctx, cb, P->species, &st, nw, &g);
                store_state(P, i, &nw);
                if (collides && nw.active) do_one_collision(ctx, mp, P, &nw, i, &g, count, &births);
                trem -= dt;
                substeps++;
            }
        }
        ctx->stats.substeps += substeps;
        ctx->stats.births += births;
        ctx->stats.rows += ilast - iup;
        total += ilast - iup;
        P->iup = ilast;
    }
    return total;
}

/* ------------------------------------------------------------------------------------ */
/* exported API (mirror of include/particulator_b200.h, prefix ora_)                     */
/* ------------------------------------------------------------------------------------ */
#define EXPORT __attribute__((visibility("default")))

EXPORT int32_t ora_abi_version(void) { return PTL_ABI_VERSION; }

EXPORT int32_t ora_context_create(int32_t device, void* stream, ora_context** out) {
    (void)device; (void)stream;
    init_constants();
    ora_context* c = calloc(1, sizeof(ora_context));
    if (!c) return PTL_ENOMEM;
    c->next_uid = 1;
    *out = c;
    return 0;
}

EXPORT int32_t ora_context_destroy(ora_context* ctx) {
    if (!ctx) return PTL_EINVAL;
    for (int i = 0; i < ctx->ntab; i++) { free(ctx->tab[i].rate); free(ctx->tab[i].ratebound); free(ctx->tab[i].lrate); free(ctx->tab[i].rbvec); }
    for (int i = 0; i < ctx->nsb; i++) { free(ctx->sb[i].log_energy); free(ctx->sb[i].data); }
    for (int i = 0; i < ctx->ncl; i++) { free(ctx->cl[i].ec); free(ctx->cl[i].pc); }
    for (int i = 0; i < ctx->npop; i++) {
        pop_t* P = &ctx->pop[i];
        free(P->x); free(P->p); free(P->w); free(P->t); free(P->s); free(P->r); free(P->active); free(P->uid);
    }
    for (int i = 0; i < PTL_MAX_WALLS; i++) { free(ctx->wall[i].x); free(ctx->wall[i].p); free(ctx->wall[i].w); free(ctx->wall[i].t); }
    free(ctx);
    return 0;
}

EXPORT const char* ora_last_error(ora_context* ctx) { return ctx ? ctx->err : "null context"; }
EXPORT int32_t ora_error_flags(ora_context* ctx, int32_t clear) { int32_t f = ctx->flags; if (clear) ctx->flags = 0; return f; }
EXPORT int32_t ora_synchronize(ora_context* ctx) { (void)ctx; return 0; }
EXPORT int32_t ora_set_rng(ora_context* ctx, uint64_t seed, uint32_t step) { ctx->seed = seed; ctx->step = step; return 0; }
EXPORT int32_t ora_get_rng(ora_context* ctx, uint64_t* seed, uint32_t* step) { *seed = ctx->seed; *step = ctx->step; return 0; }
EXPORT int32_t ora_set_uid_counter(ora_context* ctx, uint64_t next_uid) { if (!ctx || next_uid == 0) return PTL_EINVAL; ctx->next_uid = next_uid; return 0; }
EXPORT uint64_t ora_get_uid_counter(ora_context* ctx) { return ctx ? ctx->next_uid : 0; }

static double* dupd(const double* src, size_t n) {
    double* d = malloc(sizeof(double) * (n ? n : 1));
    if (src && n) memcpy(d, src, sizeof(double) * n);
    return d;
}

EXPORT int32_t ora_sb_table_create(ora_context* ctx, int32_t ncum, int32_t nE, const double* log_energy, const double* data) {
    if (ctx->nsb >= MAX_SB) return PTL_ENOMEM;
    sb_table* s = &ctx->sb[ctx->nsb];
    s->ncum = ncum; s->nE = nE;
    s->log_energy = dupd(log_energy, nE);
    s->data = dupd(data, (size_t)ncum * nE);
    return ctx->nsb++;
}

EXPORT int32_t ora_table_create_cheb(ora_context* ctx, int32_t order, int32_t nprocs, int32_t k, double xmax,
                                     const double* rate, const double* ratebound, const ptl_process_desc* procs) {
    if (ctx->ntab >= MAX_TABLES || nprocs > PTL_MAX_PROCS || order > MAX_ORDER) return PTL_EINVAL;
    table_t* T = &ctx->tab[ctx->ntab];
    memset(T, 0, sizeof(*T));
    T->kind = 0; T->order = order; T->nprocs = nprocs; T->k = k; T->xmax = xmax;
    T->rate = dupd(rate, (size_t)order * nprocs * (k + 1));
    T->ratebound = dupd(ratebound, (size_t)order * (k + 1));
    if (nprocs) memcpy(T->procs, procs, sizeof(ptl_process_desc) * nprocs);
    return ctx->ntab++;
}

EXPORT int32_t ora_table_create_linear(ora_context* ctx, int32_t grid_kind, double L1, double L2, int32_t nE, int32_t nprocs,
                                       const double* rate, double maxrate, const ptl_process_desc* procs) {
    if (ctx->ntab >= MAX_TABLES || nprocs > PTL_MAX_PROCS) return PTL_EINVAL;
    table_t* T = &ctx->tab[ctx->ntab];
    memset(T, 0, sizeof(*T));
    T->kind = 1; T->grid_kind = grid_kind; T->L1 = L1; T->L2 = L2; T->nE = nE; T->nprocs = nprocs; T->maxrate = maxrate;
    T->lrate = dupd(rate, (size_t)nprocs * nE);
    if (nprocs) memcpy(T->procs, procs, sizeof(ptl_process_desc) * nprocs);
    return ctx->ntab++;
}

EXPORT int32_t ora_table_create_linear_vb(ora_context* ctx, int32_t grid_kind, double L1, double L2, int32_t nE, int32_t nprocs,
                                          const double* rate, const double* ratebound_vec, const ptl_process_desc* procs) {
    if (!ratebound_vec) return PTL_EINVAL;
    double mx = 0;
    for (int e = 0; e < nE; e++) mx = ratebound_vec[e] > mx ? ratebound_vec[e] : mx;
    int32_t id = ora_table_create_linear(ctx, grid_kind, L1, L2, nE, nprocs, rate, mx, procs);
    if (id >= 0) ctx->tab[id].rbvec = dupd(ratebound_vec, (size_t)nE);
    return id;
}

EXPORT int32_t ora_cheb_loss_create(ora_context* ctx, int32_t order, int32_t k, double xmax, const double* ec, const double* pc) {
    if (ctx->ncl >= MAX_CHEBLOSS || order > MAX_ORDER) return PTL_EINVAL;
    cheb_loss_t* c = &ctx->cl[ctx->ncl];
    c->order = order; c->k = k; c->xmax = xmax;
    c->ec = dupd(ec, (size_t)order * (k + 1));
    c->pc = dupd(pc, (size_t)order * (k + 1));
    return ctx->ncl++;
}

EXPORT int32_t ora_table_eval(ora_context* ctx, int32_t table, int64_t n, const double* energy, double* rates_out, double* bound_out) {
    if (table < 0 || table >= ctx->ntab) return PTL_EHANDLE;
    const table_t* T = &ctx->tab[table];
    for (int64_t i = 0; i < n; i++) {
        pre_t pre;
        presample(T, energy[i], &pre);
        for (int j = 0; j < T->nprocs; j++) rates_out[j + (size_t)T->nprocs * i] = table_rate(T, j, &pre);
        bound_out[i] = table_ratebound(ctx, T, energy[i]);
    }
    return 0;
}

EXPORT int32_t ora_population_create(ora_context* ctx, int32_t species, int64_t capacity, double energy_cut, int32_t table) {
    if (ctx->npop >= MAX_POPS) return PTL_ENOMEM;
    if (table < 0 || table >= ctx->ntab) return PTL_EHANDLE;
    pop_t* P = &ctx->pop[ctx->npop];
    memset(P, 0, sizeof(*P));
    P->used = 1; P->species = species; P->capacity = capacity; P->energy_cut = energy_cut; P->table = table;
    size_t c = (size_t)(capacity > 0 ? capacity : 1);
    P->x = calloc(3 * c, sizeof(double)); P->p = calloc(3 * c, sizeof(double));
    P->w = calloc(c, sizeof(double)); P->t = calloc(c, sizeof(double));
    P->s = calloc(c, sizeof(double)); P->r = calloc(c, sizeof(double));
    P->active = calloc(c, 1); P->uid = calloc(c, sizeof(uint64_t));
    return ctx->npop++;
}

EXPORT int32_t ora_population_destroy(ora_context* ctx, int32_t pop) { (void)ctx; (void)pop; return 0; }

#define GETPOP(ctx, pop, errval) if ((pop) < 0 || (pop) >= (ctx)->npop) return (errval); pop_t* P = &(ctx)->pop[pop]

EXPORT int32_t ora_population_upload(ora_context* ctx, int32_t pop, int64_t n, const double* x3, const double* p3,
                                     const double* w, const double* t, const double* s, const double* r,
                                     const uint8_t* active, const uint64_t* uid) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    if (n > P->capacity) return PTL_EINVAL;
    memcpy(P->x, x3, sizeof(double) * 3 * n); memcpy(P->p, p3, sizeof(double) * 3 * n);
    memcpy(P->w, w, sizeof(double) * n); memcpy(P->t, t, sizeof(double) * n);
    memcpy(P->s, s, sizeof(double) * n); memcpy(P->r, r, sizeof(double) * n);
    memcpy(P->active, active, n);
    if (uid) {            /* explicit uids: the counter moves past the largest sequential one (restart safety) */
        memcpy(P->uid, uid, sizeof(uint64_t) * n);
        for (int64_t i = 0; i < n; i++)
            if (!(uid[i] & PTL_UID_HASHED_BIT) && uid[i] >= ctx->next_uid) ctx->next_uid = uid[i] + 1;
    } else {
        for (int64_t i = 0; i < n; i++) P->uid[i] = ctx->next_uid + (uint64_t)i;
        ctx->next_uid += (uint64_t)n;
    }
    P->n = n; P->iup = 0;
    return 0;
}

EXPORT int64_t ora_population_download(ora_context* ctx, int32_t pop, int64_t max_n, double* x3, double* p3, double* w,
                                       double* t, double* s, double* r, uint8_t* active, uint64_t* uid) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    int64_t n = P->n < max_n ? P->n : max_n;
    if (x3) memcpy(x3, P->x, sizeof(double) * 3 * n);
    if (p3) memcpy(p3, P->p, sizeof(double) * 3 * n);
    if (w) memcpy(w, P->w, sizeof(double) * n);
    if (t) memcpy(t, P->t, sizeof(double) * n);
    if (s) memcpy(s, P->s, sizeof(double) * n);
    if (r) memcpy(r, P->r, sizeof(double) * n);
    if (active) memcpy(active, P->active, n);
    if (uid) memcpy(uid, P->uid, sizeof(uint64_t) * n);
    return n;
}

EXPORT int64_t ora_population_n(ora_context* ctx, int32_t pop) { GETPOP(ctx, pop, PTL_EHANDLE); return P->n; }
EXPORT int64_t ora_population_capacity(ora_context* ctx, int32_t pop) { GETPOP(ctx, pop, PTL_EHANDLE); return P->capacity; }
EXPORT int32_t ora_population_clear(ora_context* ctx, int32_t pop) { GETPOP(ctx, pop, PTL_EHANDLE); P->n = 0; return 0; }
EXPORT int32_t ora_population_set_n(ora_context* ctx, int32_t pop, int64_t n) { GETPOP(ctx, pop, PTL_EHANDLE); P->n = n; return 0; }
EXPORT void* ora_population_column_ptr(ora_context* ctx, int32_t pop, int32_t col) { (void)ctx; (void)pop; (void)col; return NULL; }

EXPORT int64_t ora_population_append(ora_context* ctx, int32_t pop, const double* x3, const double* p3, double w, double t,
                                     double s, double r, uint64_t uid) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    state_t st;
    for (int c = 0; c < 3; c++) { st.x.v[c] = x3[c]; st.p.v[c] = p3[c]; }
    st.w = w; st.t = t; st.s = s; st.r = r; st.active = 1;
    if (uid == 0) uid = ctx->next_uid++;
    else if (!(uid & PTL_UID_HASHED_BIT) && uid >= ctx->next_uid) ctx->next_uid = uid + 1;
    if (P->n >= P->capacity && kinenergy(P->species, st.p) > P->energy_cut) { ctx->flags |= PTL_ERR_CAPACITY_OVERFLOW; return PTL_ECAPACITY; }
    return add_particle(ctx, P, &st, uid);
}

EXPORT int32_t ora_population_deactivate(ora_context* ctx, int32_t pop, int64_t i) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    if (i < 0 || i >= P->n) return PTL_EINVAL;
    P->active[i] = 0;
    return 0;
}

static void copy_row(pop_t* P, int64_t dst, int64_t src) {
    for (int c = 0; c < 3; c++) { P->x[3 * dst + c] = P->x[3 * src + c]; P->p[3 * dst + c] = P->p[3 * src + c]; }
    P->w[dst] = P->w[src]; P->t[dst] = P->t[src]; P->s[dst] = P->s[src]; P->r[dst] = P->r[src];
    P->active[dst] = P->active[src]; P->uid[dst] = P->uid[src];
}

/* repack!: population.jl:229-259 (1-based l, i in the reference; 0-based rows here) */
EXPORT int64_t ora_repack(ora_context* ctx, int32_t pop) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    int64_t l = P->n; /* 1-based index of last candidate */
    if (l == 0) return 0;
    while (l > 0 && !P->active[l - 1]) l--;
    if (l == 0) { P->n = 0; return 0; }
    int64_t i = 1;
    while (i <= l) {
        if (!P->active[i - 1]) {
            copy_row(P, i - 1, l - 1);
            l--;
            while (!P->active[l - 1]) l--;
        }
        i++;
    }
    P->n = l;
    return l;
}

/* droplow!: population.jl:273-284 */
EXPORT int64_t ora_droplow(ora_context* ctx, int32_t pop, double thres) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    double th = thres == 0 ? P->energy_cut : thres;
    for (int64_t i = 0; i < P->n; i++) {
        if (P->active[i]) {
            vec3 p = {{P->p[3 * i], P->p[3 * i + 1], P->p[3 * i + 2]}};
            if (kinenergy(P->species, p) < th) P->active[i] = 0;
        }
    }
    return ora_repack(ctx, pop);
}

/* diagnostics: population.jl:78-223 (serial sums in row order) */
EXPORT int32_t ora_diag(ora_context* ctx, int32_t pop, ptl_diag_out* out) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    memset(out, 0, sizeof(*out));
    out->n = P->n;
    double maxe = -INFINITY;
    for (int64_t i = 0; i < P->n; i++) {
        vec3 p = {{P->p[3 * i], P->p[3 * i + 1], P->p[3 * i + 2]}};
        double e = kinenergy(P->species, p);
        if (e > maxe) maxe = e;
        if (!P->active[i]) continue;
        double w = P->w[i];
        out->nactive++;
        out->weight += w;
        out->wenergy += w * e;
        double d = 0;
        for (int c = 0; c < 3; c++) {
            double xc = P->x[3 * i + c];
            out->wx[c] += w * xc;
            out->wx2[c] += w * xc * xc;
            d += xc * xc;
        }
        out->wr2 += w * d;
    }
    out->maxenergy = maxe;
    return 0;
}

EXPORT int32_t ora_histogram(ora_context* ctx, int32_t pop, int32_t quantity, double lo, double hi, int32_t nbins,
                             int32_t logscale, double* out) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    for (int b = 0; b < nbins; b++) out[b] = 0;
    double a = logscale ? log10(lo) : lo, bb = logscale ? log10(hi) : hi;
    for (int64_t i = 0; i < P->n; i++) {
        if (!P->active[i]) continue;
        vec3 p = {{P->p[3 * i], P->p[3 * i + 1], P->p[3 * i + 2]}};
        double q = quantity == 0 ? kinenergy(P->species, p) : p.v[2] / norm3(p);
        if (logscale) { if (!(q > 0)) continue; q = log10(q); }
        double f = (q - a) / (bb - a) * nbins;
        if (!(f >= 0) || !(f < nbins)) continue;
        out[(int)f] += P->w[i];
    }
    return 0;
}

/* roulette!: population.jl:291-309 (constant p) */
/* energy-dependent law of roulette!(f, popl) / split!(f, popl): f tabulated on n nodes uniform in E or log10(E) between lo and hi,
 * linear interpolation, flat outside (n == 1: constant) */
static double law_value(double eng, double lo, double hi, int32_t n, int32_t logscale, const double* v) {
    if (n <= 1) return v[0];
    double x = logscale ? log10(eng) : eng;
    double u = (x - lo) / (hi - lo) * (double)(n - 1);
    if (!(u > 0)) return v[0];
    if (u >= (double)(n - 1)) return v[n - 1];
    int k = (int)u;
    double f = u - (double)k;
    return v[k] * (1 - f) + v[k + 1] * f;
}

/* roulette!(f, popl): population.jl:291-309 */
EXPORT int32_t ora_roulette_law(ora_context* ctx, int32_t pop, double lo, double hi, int32_t n, int32_t logscale, const double* pv) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    if (n < 1 || !pv || (n > 1 && !(hi > lo))) return PTL_EINVAL;
    for (int64_t i = 0; i < P->n; i++) {
        if (!P->active[i]) continue;
        double p = law_value(kinenergy(P->species, load_state(P, i).p), lo, hi, n, logscale, pv);
        rng_t g;
        rng_init(&g, P->uid[i], DOM_ROULETTE, ctx->seed, ctx->step);
        if (rng_u(&g) < p) P->w[i] /= p;
        else P->active[i] = 0;
    }
    ctx->step++;
    return 0;
}
EXPORT int32_t ora_roulette(ora_context* ctx, int32_t pop, double p) { return ora_roulette_law(ctx, pop, 0, 1, 1, 0, &p); }

/* split!(f, popl): population.jl:316-335; Poisson(p) by sequential inversion; copies get
 * fresh uids so that their streams differ (the reference relies on a shared global RNG) */
EXPORT int32_t ora_split_law(ora_context* ctx, int32_t pop, double lo, double hi, int32_t n, int32_t logscale, const double* pv) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    if (n < 1 || !pv || (n > 1 && !(hi > lo))) return PTL_EINVAL;
    int64_t n0 = P->n;
    for (int64_t i = 0; i < n0; i++) {
        if (!P->active[i]) continue;
        double p = law_value(kinenergy(P->species, load_state(P, i).p), lo, hi, n, logscale, pv);
        P->w[i] /= (1 + p);
        rng_t g;
        rng_init(&g, P->uid[i], DOM_SPLIT, ctx->seed, ctx->step);
        double u = rng_u(&g);
        double pk = exp(-p), cdf = pk;
        int k = 0;
        while (u > cdf && k < 1000) { k++; pk *= p / k; cdf += pk; }
        state_t st = load_state(P, i);
        for (int c = 0; c < k; c++) {
            uint64_t cu[2];
            child_uids(P->uid[i] ^ ((uint64_t)DOM_SPLIT << 32), (uint32_t)c, ctx->seed, ctx->step, cu);
            add_particle(ctx, P, &st, cu[0]);
        }
    }
    ctx->step++;
    return ctx->flags;
}
EXPORT int32_t ora_split(ora_context* ctx, int32_t pop, double p) { return ora_split_law(ctx, pop, 0, 1, 1, 0, &p); }

/* shuffle!(popl): population.jl:266-271, a uniformly distributed permutation of the rows.  The reference runs Fisher-Yates
 * on the global RNG; here every row draws a 64-bit key from the Philox stream of (row index, seed, step) and the rows are
 * sorted by key (ties by row index): also uniform over permutations, and reproducible on any number of threads. */
typedef struct { uint64_t key; int64_t row; } shuf_t;
static int shuf_cmp(const void* a, const void* b) {
    const shuf_t *x = a, *y = b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}
EXPORT int32_t ora_shuffle(ora_context* ctx, int32_t pop) {
    GETPOP(ctx, pop, PTL_EHANDLE);
    int64_t n = P->n;
    if (n > 1) {
        shuf_t* k = malloc(sizeof(shuf_t) * (size_t)n);
        for (int64_t i = 0; i < n; i++) {
            uint32_t key[2] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32) ^ DOM_SHUFFLE};
            uint32_t ctr[4] = {0, ctx->step, (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)}, o[4];
            philox4x32_10(ctr, key, o);
            k[i].key = ((uint64_t)o[1] << 32) | o[0];
            k[i].row = i;
        }
        qsort(k, (size_t)n, sizeof(shuf_t), shuf_cmp);
        state_t* tmp = malloc(sizeof(state_t) * (size_t)n);
        uint64_t* tu = malloc(sizeof(uint64_t) * (size_t)n);
        for (int64_t i = 0; i < n; i++) { tmp[i] = load_state(P, k[i].row); tu[i] = P->uid[k[i].row]; }
        for (int64_t i = 0; i < n; i++) { store_state(P, i, &tmp[i]); P->uid[i] = tu[i]; }
        free(tmp); free(tu); free(k);
    }
    ctx->step++;
    return 0;
}

EXPORT int32_t ora_multipop_create(ora_context* ctx, const int32_t* pops, int32_t count) {
    if (ctx->nmp >= MAX_MP || count > PTL_NSPECIES) return PTL_EINVAL;
    multipop_t* m = &ctx->mp[ctx->nmp];
    m->npop = count;
    for (int s = 0; s < PTL_NSPECIES; s++) m->by_species[s] = -1;
    for (int i = 0; i < count; i++) {
        if (pops[i] < 0 || pops[i] >= ctx->npop) return PTL_EHANDLE;
        m->pops[i] = pops[i];
        m->by_species[ctx->pop[pops[i]].species] = pops[i];
    }
    return ctx->nmp++;
}

/* init!: mixed_population.jl:20-35 ; also advance_init!: :97-110 */
static void init_r(ora_context* ctx, const multipop_t* mp, int reset_iup) {
    for (int ip = 0; ip < mp->npop; ip++) {
        pop_t* P = &ctx->pop[mp->pops[ip]];
        if (reset_iup) P->iup = 0;
#pragma omp parallel for schedule(static)
        for (int64_t i = P->iup; i < P->n; i++) {
            if (!P->active[i]) continue;
            vec3 p = {{P->p[3 * i], P->p[3 * i + 1], P->p[3 * i + 2]}};
            P->r[i] = setr_value(ctx, P, p);
        }
    }
}

EXPORT int32_t ora_init(ora_context* ctx, int32_t mp) {
    if (mp < 0 || mp >= ctx->nmp) return PTL_EHANDLE;
    init_r(ctx, &ctx->mp[mp], 0);
    return ctx->flags;
}

/* advance!: mixed_population.jl:38-47 */
EXPORT int32_t ora_advance(ora_context* ctx, int32_t mp, const ptl_pusher_desc* pusher, double tfinal, const ptl_callback_desc* cb) {
    if (mp < 0 || mp >= ctx->nmp) return PTL_EHANDLE;
    const multipop_t* M = &ctx->mp[mp];
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    init_r(ctx, M, 1);
    int64_t n = 1;
    while (n > 0) {
        n = advance1(ctx, M, pusher, tfinal, cb);
        ctx->stats.passes++;
    }
    ctx->step++;
    return ctx->flags;
}

EXPORT int32_t ora_set_profiling(ora_context* ctx, int32_t on) { (void)ctx; (void)on; return 0; }
EXPORT int64_t ora_launch_count(ora_context* ctx, int32_t reset) { (void)ctx; (void)reset; return 0; }
EXPORT int32_t ora_last_advance_stats(ora_context* ctx, ptl_advance_stats* out) { *out = ctx->stats; return 0; }

EXPORT int32_t ora_collision_counts(ora_context* ctx, int32_t table, int64_t* counts, int32_t clear) {
    if (table < 0 || table >= ctx->ntab) return PTL_EHANDLE;
    table_t* T = &ctx->tab[table];
    for (int j = 0; j <= T->nprocs; j++) { counts[j] = T->counts[j]; if (clear) T->counts[j] = 0; }
    return 0;
}

EXPORT int64_t ora_wall_records(ora_context* ctx, int32_t iwall, int64_t max_n, double* x3, double* p3, double* w, double* t,
                                int32_t clear) {
    if (iwall < 0 || iwall >= PTL_MAX_WALLS) return PTL_EINVAL;
    wallrec_t* W = &ctx->wall[iwall];
    int64_t n = W->n < max_n ? W->n : max_n;
    if (x3) memcpy(x3, W->x, sizeof(double) * 3 * n);
    if (p3) memcpy(p3, W->p, sizeof(double) * 3 * n);
    if (w) memcpy(w, W->w, sizeof(double) * n);
    if (t) memcpy(t, W->t, sizeof(double) * n);
    int64_t total = W->n;
    if (clear) W->n = 0;
    return total;
}

/* Test hook: run collide() of process `j` of `table` once per input momentum (row i uses uid
 * uid0 + i, draw stream from index 0) and report the outcome without touching any population.
 * out[i*24 ..]: [0]=outcome kind, [1]=sp2, [2]=sp3, [3]=draws consumed, [4..6]=p1, [7]=s1,
 * [8..10]=p2, [11]=s2, [12..14]=p3, [15]=s3, [16..23] reserved. */
EXPORT int32_t ora_collide_test(ora_context* ctx, int32_t species, int32_t table, int32_t j, int64_t n, const double* p3,
                                uint64_t uid0, double* out) {
    if (table < 0 || table >= ctx->ntab) return PTL_EHANDLE;
    const table_t* T = &ctx->tab[table];
    if (j < 0 || j >= T->nprocs) return PTL_EINVAL;
    for (int64_t i = 0; i < n; i++) {
        rng_t g;
        rng_init(&g, uid0 + (uint64_t)i, DOM_COLLISION, ctx->seed, ctx->step);
        state_t st;
        memset(&st, 0, sizeof(st));
        for (int c = 0; c < 3; c++) st.p.v[c] = p3[3 * i + c];
        st.w = 1.0; st.active = 1;
        outcome_t o;
        memset(&o, 0, sizeof(o));
        collide(ctx, &g, &T->procs[j], species, &st, kinenergy(species, st.p), &o);
        double* r = out + 24 * i;
        memset(r, 0, sizeof(double) * 24);
        r[0] = o.kind; r[1] = o.sp2; r[2] = o.sp3; r[3] = g.idx;
        if (o.kind == OUT_STATE_CHANGE || o.kind == OUT_NEW_PARTICLE) { for (int c = 0; c < 3; c++) r[4 + c] = o.s1.p.v[c]; r[7] = o.s1.s; }
        if (o.kind == OUT_NEW_PARTICLE || o.kind == OUT_REPLACE || o.kind == OUT_REPLACE_PAIR) { for (int c = 0; c < 3; c++) r[8 + c] = o.s2.p.v[c]; r[11] = o.s2.s; }
        if (o.kind == OUT_REPLACE_PAIR) { for (int c = 0; c < 3; c++) r[12 + c] = o.s3.p.v[c]; r[15] = o.s3.s; }
    }
    return 0;
}

/* Replay of REFERENCE-emitted vectors (julia/emit_golden.jl -> tests/golden/reference_vectors.npz): collide() of process j with
 * the reference's own rand() sequence injected draw by draw (uniforms[nu] per event).  Same output layout as ora_collide_test;
 * out[3] = draws consumed.  This is what pins the samplers against the reference itself once a Julia runtime is available. */
EXPORT int32_t ora_collide_replay(ora_context* ctx, int32_t species, int32_t table, int32_t j, int64_t n, const double* p3,
                                  const double* uniforms, int32_t nu, double* out) {
    if (table < 0 || table >= ctx->ntab) return PTL_EHANDLE;
    const table_t* T = &ctx->tab[table];
    if (j < 0 || j >= T->nprocs || nu < 1 || !uniforms) return PTL_EINVAL;
    for (int64_t i = 0; i < n; i++) {
        rng_t g;
        rng_init(&g, 1, DOM_COLLISION, 0, 0);
        g.inject = uniforms + (size_t)nu * i;
        g.ninject = (uint32_t)nu;
        state_t st;
        memset(&st, 0, sizeof(st));
        for (int c = 0; c < 3; c++) st.p.v[c] = p3[3 * i + c];
        st.w = 1.0; st.active = 1;
        outcome_t o;
        memset(&o, 0, sizeof(o));
        collide(ctx, &g, &T->procs[j], species, &st, kinenergy(species, st.p), &o);
        double* r = out + 24 * i;
        memset(r, 0, sizeof(double) * 24);
        r[0] = o.kind; r[1] = o.sp2; r[2] = o.sp3; r[3] = g.idx;
        if (o.kind == OUT_STATE_CHANGE || o.kind == OUT_NEW_PARTICLE) { for (int c = 0; c < 3; c++) r[4 + c] = o.s1.p.v[c]; r[7] = o.s1.s; }
        if (o.kind == OUT_NEW_PARTICLE || o.kind == OUT_REPLACE || o.kind == OUT_REPLACE_PAIR) { for (int c = 0; c < 3; c++) r[8 + c] = o.s2.p.v[c]; r[11] = o.s2.s; }
        if (o.kind == OUT_REPLACE_PAIR) { for (int c = 0; c < 3; c++) r[12 + c] = o.s3.p.v[c]; r[15] = o.s3.s; }
    }
    return 0;
}

/* Test hook: n uniforms of the stream (uid, domain 0, seed, step) starting at draw index 0, and
 * the raw Philox block for KAT checks. */
EXPORT int32_t ora_rng_test(ora_context* ctx, uint64_t uid, uint64_t seed, uint32_t step, int32_t n, double* out) {
    (void)ctx;
    rng_t g;
    rng_init(&g, uid, DOM_COLLISION, seed, step);
    for (int i = 0; i < n; i++) out[i] = rng_u(&g);
    return 0;
}
EXPORT int32_t ora_philox_test(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    philox4x32_10(ctr, key, out);
    return 0;
}
/* torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline asks for all host cores explicitly */
EXPORT int32_t ora_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
    return 0;
}
EXPORT int32_t ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
