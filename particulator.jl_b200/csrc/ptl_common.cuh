// ptl_common.cuh — device-side views, constants and the counter-based RNG shared by all kernels.
//
// Data layout in HBM: every population is a planar SoA of 10 double columns
// (x0,x1,x2,p0,p1,p2,w,t,s,r), one uint8 `active` column and one uint64 `uid` column, each
// `capacity` long and 256-byte aligned, so that a warp reading one column touches one
// contiguous 256 B (double) span -> 100 % sector efficiency.  The particle count `n` of each
// population lives in device memory (births append with atomicAdd).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/particulator_b200.h"

// Code placement.  ptxas lays the out-of-line device functions of a kernel out behind its body in the order of their
// mangled names.  -DPTL_HOT_NAMES renames the four functions the hot work units call so that they sort first, i.e. sit
// right behind the kernel body instead of up to 150 KB away from it (instruction-cache experiment).
#ifdef PTL_HOT_NAMES
#define philox_block a_hot_phlx
#define nlog a_hot_nlog
#define nsincospi a_hot_scpi
#if PTL_HOT_NAMES > 1      // (add_particle is 615 instructions and only runs for births above the cut: it is NOT hot; =2 reproduces the first measurement)
#define add_particle a_hot_addp
#endif
#endif

namespace ptl {

// ---- constants: reference src/constants.jl (CODATA-2014), same expression order ----------------
constexpr double CO_C = 299792458.0;                  // constants.jl:37
constexpr double CO_E = 1.6021766208e-19;             // constants.jl:50-54
constexpr double CO_ME = 9.10938356e-31;              // constants.jl:52
constexpr double CO_EPS0 = 8.854187817620389e-12;     // constants.jl:55
constexpr double CO_ALPHA = 0.0072973525664;          // constants.jl:60
constexpr double CO_HBAR = 1.0545718001391127e-34;    // constants.jl:77
constexpr double CO_PI = 3.141592653589793;
constexpr double CO_RE = (CO_E * CO_E) / (CO_ME * (CO_C * CO_C)) / (4 * CO_PI * CO_EPS0);  // :157
constexpr double CO_A0 = CO_HBAR / (CO_ME * CO_C * CO_ALPHA);                               // :160
constexpr double CO_MC2 = CO_ME * (CO_C * CO_C);                                            // :163
constexpr double CO_C2 = CO_C * CO_C;
constexpr double INV_MC2 = 1.0 / CO_MC2;
constexpr double INV_C = 1.0 / CO_C;
constexpr double INV_ME = 1.0 / CO_ME;
constexpr double C2_OVER_MC2SQ = CO_C2 / (CO_MC2 * CO_MC2);
constexpr double DBL_EPS = 2.220446049250313e-16;     // eps(Float64), mixed_population.jl:66

constexpr int MAX_ORDER = 8;
constexpr int MAX_SB = 8;
constexpr int MAX_CHEBLOSS = 4;

enum Column { COL_X0 = 0, COL_X1, COL_X2, COL_P0, COL_P1, COL_P2, COL_W, COL_T, COL_S, COL_R, COL_ACTIVE, COL_UID, NCOLS };

struct PopView {
    double* col[10];                 // x0,x1,x2,p0,p1,p2,w,t,s,r
    uint8_t* active;
    uint64_t* uid;
    unsigned long long* n;           // device-resident particle count (population.jl:9)
    long long capacity;
    double energy_cut;
    int species;
    int present;
};

struct TableView {
    int kind;                        // 0 = Chebyshev, 1 = linear
    int nprocs;
    int order, k;
    double xmax;
    double rxmax;                    // RN(1 / xmax), host-computed: precheb divides by Markstein's exact sequence
    const double* rate;              // cheb: [order, nprocs, k+1]; linear: [nprocs, nE]
    const double* ratebound;         // cheb: [order, k+1]
    const double* cum;               // running sum over processes of `rate` (same layout): selection accelerator
    const double2* cum2;             // linear tables: {cum[j, e], cum[j, e + 1]} per (j, e): one 16-byte load per search probe
    int grid_kind, nE;
    double L1, L2, maxrate;
    const double* rbvec;             // linear tables: vector rate bound on the energy grid (collision_table.jl:35-43) or nullptr
    const ptl_process_desc* procs;   // device copy
    unsigned long long* counts;      // [nprocs + 1]
    unsigned long long mono_mask;    // cheb, order 3: bit i set <=> every fitted rate is >= 0 on interval i, so the running
                                     // sums over processes are non-decreasing there (binary-search selection is valid)
};

struct SbView {
    int ncum, nE;
    const double* log_energy;
    const double* data;              // [ncum, nE], ncum fastest
};

struct ChebLossView {
    int order, k;
    double xmax;
    double rxmax;
    const double* ec;
    const double* pc;
};

struct WallBuf {
    double* col[8];                  // x0,x1,x2,p0,p1,p2,w,t
    unsigned long long* n;
    long long capacity;
};

struct AdvanceParams {
    PopView pop[PTL_NSPECIES];       // indexed by species (MultiPopulation index, mixed_population.jl:15)
    TableView tab[PTL_NSPECIES];     // table of that species' population
    SbView sb[MAX_SB];
    ChebLossView cl[MAX_CHEBLOSS];
    ptl_pusher_desc pusher;
    ptl_callback_desc cb;
    WallBuf wall[PTL_MAX_WALLS];
    double tfinal;
    uint32_t seed_lo, seed_hi, step;
    int has_cb;
    int fast_force;                  // 1: the pusher is RK2 with one homogeneous E field and no B for every species (hot path)
    double fastE[3];                 // e * E  [N]; the species' charge sign / mass is applied in the kernel
    int* flags;                      // sticky PTL_ERR_* bits
    unsigned long long* substeps;    // sub-step counters, one per species
    unsigned long long* dbg;         // PTL_TRACE scheduler statistics (max rounds, sum rounds, CTAs)
    unsigned long long* births;
};

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr uint32_t DOM_COLLISION = 0u;
constexpr uint32_t DOM_CHILD_UID = 0x5EED0001u;
constexpr uint32_t DOM_ROULETTE = 0x5EED0002u;
constexpr uint32_t DOM_SPLIT = 0x5EED0003u;
constexpr uint32_t DOM_SHUFFLE = 0x5EED0004u;

// 64 random bits -> double in the OPEN interval (0,1): (m + 0.5) * 2^-52 with m the top 52 bits.
// Built as [1,2) mantissa + one exact add (no int->double conversion).
__device__ __forceinline__ double bits_to_u01(uint32_t lo, uint32_t hi) {
    uint32_t dhi = 0x3FF00000u | (hi >> 12);
    uint32_t dlo = (hi << 20) | (lo >> 12);
    return __hiloint2double((int)dhi, (int)dlo) + (-1.0 + 0x1.0p-53);
}

// One Philox block as a real function: the advance kernels draw at ~70 sites and inlining the ten
// rounds everywhere blew the kernel up to 300 KB of SASS (instruction-cache misses were the top stall).
static __device__ __noinline__ uint4 philox_block(uint32_t block, uint32_t step, uint32_t seed_lo, uint32_t seed_hi, uint32_t k0, uint32_t k1) {
    uint32_t o[4];
    philox4x32_10(block, step, seed_lo, seed_hi, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// Per-particle stream: replaces the task-local rand() of the reference (src/util.jl:17 and every
// collide).  The n-th uniform of (uid, advance call) is word pair (n & 1) of block n >> 1.
struct Rng {
    uint32_t k0, k1;
    uint32_t idx;
    uint32_t cblock;
    uint32_t c2, c3;

    __device__ __forceinline__ void init(uint64_t uid, uint32_t domain) {
        k0 = (uint32_t)uid;
        k1 = (uint32_t)(uid >> 32) ^ domain;
        idx = 0;
        cblock = 0xFFFFFFFFu;
        c2 = c3 = 0;
    }
    __device__ __forceinline__ double u(uint32_t step, uint32_t seed_lo, uint32_t seed_hi) {
        uint32_t block = idx >> 1;
        uint32_t lo, hi;
        if ((idx & 1) && cblock == block) {
            lo = c2; hi = c3;
        } else {
            uint4 o = philox_block(block, step, seed_lo, seed_hi, k0, k1);
            c2 = o.z; c3 = o.w; cblock = block;
            lo = (idx & 1) ? o.z : o.x;
            hi = (idx & 1) ? o.w : o.y;
        }
        idx++;
        return bits_to_u01(lo, hi);
    }
    __device__ __forceinline__ void skip() { idx++; }
};

__device__ __forceinline__ void child_uids(uint64_t parent, uint32_t idx, uint32_t step, uint32_t seed_lo, uint32_t seed_hi,
                                           uint64_t out[2]) {
    uint4 o = philox_block(idx, step, seed_lo, seed_hi, (uint32_t)parent, (uint32_t)(parent >> 32) ^ DOM_CHILD_UID);
    // uid space: bit 63 set = hashed (births), clear = sequential (host-assigned); include/particulator_b200.h
    out[0] = (((uint64_t)o.y << 32) | o.x) | PTL_UID_HASHED_BIT;
    out[1] = (((uint64_t)o.w << 32) | o.z) | PTL_UID_HASHED_BIT;
}

}  // namespace ptl
