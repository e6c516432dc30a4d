// ptl_store.cuh — kernels of the device-resident particle store: layout transposes for
// upload/download, K2 droplow!/repack! (ballot + prefix-sum stream compaction that reproduces the
// reference's tail-fill permutation), K3 fused diagnostics, histograms, K4 roulette!/split!.
#pragma once
#include "ptl_physics.cuh"

namespace ptl {

// ---- host layout (xyz interleaved, StructArray of SVector{3}: population.jl:14,37) <-> planar SoA ------
__global__ void k_aos3_to_planar(const double* __restrict__ a3, double* __restrict__ c0, double* __restrict__ c1,
                                 double* __restrict__ c2, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    c0[i] = a3[3 * i]; c1[i] = a3[3 * i + 1]; c2[i] = a3[3 * i + 2];
}
__global__ void k_planar_to_aos3(const double* __restrict__ c0, const double* __restrict__ c1, const double* __restrict__ c2,
                                 double* __restrict__ a3, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    a3[3 * i] = c0[i]; a3[3 * i + 1] = c1[i]; a3[3 * i + 2] = c2[i];
}
__global__ void k_fill_uid(uint64_t* __restrict__ uid, unsigned long long base, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) uid[i] = base + (unsigned long long)i;
}

// largest sequential uid (bit 63 clear) of an uploaded uid column -> *out (atomicMax): keeps the context's default-uid
// counter ahead of every live sequential uid without a host pass over the array
static __global__ void k_uid_max(const uint64_t* __restrict__ uid, long long n, unsigned long long* out) {
    unsigned long long m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long u = uid[i];
        if (!(u & PTL_UID_HASHED_BIT) && u > m) m = u;
    }
    for (int off = 16; off > 0; off >>= 1) { unsigned long long o = __shfl_down_sync(0xffffffffu, m, off); m = o > m ? o : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// OR a bit-set into the sticky flags word from the host side (ptl_population_append on a full population)
static __global__ void k_or_flags(int* flags, int bits) { atomicOr(flags, bits); }

// ---- K2: droplow! (population.jl:273-284) + repack! (:229-259) ---------------------------------------------
// repack! is a serial tail-fill: holes of the prefix are filled, in ascending order, by the actives of the
// tail taken in descending order.  Parallel form: L' = #actives; holes h_1<h_2<.. among rows
// [0,L'), tail actives a_1>a_2>.. among rows [L',n); move a_m -> h_m.  Same permutation, bit for bit.
constexpr int CMP_THREADS = 256;
constexpr int CMP_ROWS = 4;                       // rows per thread
constexpr int CMP_TILE = CMP_THREADS * CMP_ROWS;  // rows per block

// pass A: optional E < thres flagging + per-tile active counts (warp ballot + popc)
template <int SP>
__global__ void __launch_bounds__(CMP_THREADS) k_flag_count(PopView Q, long long n, double thres, int do_flag,
                                                            unsigned int* __restrict__ tile_counts) {
    __shared__ unsigned int wsum[CMP_THREADS / 32];
    long long base = (long long)blockIdx.x * CMP_TILE;
    unsigned int cnt = 0;
#pragma unroll
    for (int q = 0; q < CMP_ROWS; q++) {
        long long i = base + q * CMP_THREADS + threadIdx.x;
        bool a = false;
        if (i < n) {
            a = Q.active[i] != 0;
            if (a && do_flag) {
                Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
                if (kinenergy<SP>(p) < thres) {   // strict <  (population.jl:278)
                    a = false;
                    Q.active[i] = 0;
                }
            }
        }
        cnt += __popc(__ballot_sync(0xffffffffu, a));
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int s = 0;
        for (int wq = 0; wq < CMP_THREADS / 32; wq++) s += wsum[wq];
        tile_counts[blockIdx.x] = s;
    }
}

// pass B: exclusive scan of the tile counts (single block), total -> *total_out
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned int* __restrict__ tile_counts, long long ntiles,
                                                     unsigned long long* __restrict__ tile_offsets, unsigned long long* total_out) {
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = 0; base < ntiles; base += 1024) {
        long long i = base + threadIdx.x;
        unsigned long long v = i < ntiles ? tile_counts[i] : 0ULL;
        unsigned long long inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long wv = wtot[lane], winc = wv;
            for (int off = 1; off < 32; off <<= 1) {
                unsigned long long o = __shfl_up_sync(0xffffffffu, winc, off);
                if (lane >= off) winc += o;
            }
            wtot[lane] = winc - wv;   // exclusive per-warp offsets
        }
        __syncthreads();
        unsigned long long excl = carry + wtot[wid] + (inc - v);
        if (i < ntiles) tile_offsets[i] = excl;
        __syncthreads();              // everyone has read `carry`
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// pass C: emit hole list (ascending) and tail-active list (descending)
__global__ void __launch_bounds__(CMP_THREADS) k_emit_moves(PopView Q, long long n, const unsigned long long* __restrict__ tile_offsets,
                                                            const unsigned long long* __restrict__ total_ptr,
                                                            long long* __restrict__ holes, long long* __restrict__ tails,
                                                            unsigned long long* __restrict__ nmoves) {
    __shared__ unsigned int wcnt[CMP_ROWS][CMP_THREADS / 32];
    long long base = (long long)blockIdx.x * CMP_TILE;
    const long long total = (long long)*total_ptr;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool a[CMP_ROWS];
    unsigned int bal[CMP_ROWS];
#pragma unroll
    for (int q = 0; q < CMP_ROWS; q++) {
        long long i = base + q * CMP_THREADS + threadIdx.x;
        a[q] = (i < n) && Q.active[i] != 0;
        bal[q] = __ballot_sync(0xffffffffu, a[q]);
        if (lane == 0) wcnt[q][wid] = __popc(bal[q]);
    }
    __syncthreads();
    unsigned long long run = tile_offsets[blockIdx.x];
    unsigned int myholes = 0;
#pragma unroll
    for (int q = 0; q < CMP_ROWS; q++) {
        unsigned long long before = run;
        for (int wq = 0; wq < CMP_THREADS / 32; wq++) {
            if (wq < wid) before += wcnt[q][wq];
            run += wcnt[q][wq];
        }
        before += __popc(bal[q] & ((1u << lane) - 1));
        long long i = base + q * CMP_THREADS + threadIdx.x;
        if (i < n) {
            if (i < total && !a[q]) { holes[i - (long long)before] = i; myholes++; }
            if (i >= total && a[q]) tails[total - (long long)before - 1] = i;
        }
    }
    for (int off = 16; off > 0; off >>= 1) myholes += __shfl_down_sync(0xffffffffu, myholes, off);
    if (lane == 0 && myholes) atomicAdd(nmoves, (unsigned long long)myholes);
}

// pass D: move tail actives into the holes (all 12 columns)
__global__ void k_apply_moves(PopView Q, const long long* __restrict__ holes, const long long* __restrict__ tails,
                              const unsigned long long* __restrict__ nmoves) {
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= (long long)*nmoves) return;
    long long dst = holes[m], src = tails[m];
#pragma unroll
    for (int c = 0; c < 10; c++) Q.col[c][dst] = Q.col[c][src];
    Q.active[dst] = Q.active[src];
    Q.uid[dst] = Q.uid[src];
}

__global__ void k_set_count(unsigned long long* n, const unsigned long long* total) { *n = *total; }

// ---- K3: fused diagnostics (population.jl:78-223) -----------------------------------------------------------
constexpr int DIAG_NVAL = 12;   // nactive, weight, wenergy, wx[3], wx2[3], wr2, maxenergy, (pad)
constexpr int DIAG_THREADS = 256;

template <int SP>
__global__ void __launch_bounds__(DIAG_THREADS) k_diag_partial(PopView Q, long long n, double* __restrict__ partial) {
    __shared__ double sh[DIAG_NVAL][DIAG_THREADS / 32];
    double v[DIAG_NVAL];
#pragma unroll
    for (int q = 0; q < DIAG_NVAL; q++) v[q] = 0;
    v[10] = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
        double e = kinenergy<SP>(p);
        v[10] = fmax(v[10], e);   // maxenergy runs over every row < n, active or not (population.jl:172-174)
        if (!Q.active[i]) continue;
        double w = Q.col[COL_W][i];
        Vec3 x = {Q.col[COL_X0][i], Q.col[COL_X1][i], Q.col[COL_X2][i]};
        v[0] += 1; v[1] += w; v[2] += w * e;
        v[3] += w * x.x; v[4] += w * x.y; v[5] += w * x.z;
        v[6] += w * x.x * x.x; v[7] += w * x.y * x.y; v[8] += w * x.z * x.z;
        v[9] += w * dot(x, x);
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < DIAG_NVAL; q++) {
        double a = v[q];
        for (int off = 16; off > 0; off >>= 1) {
            double o = __shfl_down_sync(0xffffffffu, a, off);
            a = (q == 10) ? fmax(a, o) : a + o;
        }
        if (lane == 0) sh[q][wid] = a;
    }
    __syncthreads();
    if (threadIdx.x < DIAG_NVAL) {
        int q = threadIdx.x;
        double a = sh[q][0];
        for (int wq = 1; wq < DIAG_THREADS / 32; wq++) a = (q == 10) ? fmax(a, sh[q][wq]) : a + sh[q][wq];
        partial[(size_t)blockIdx.x * DIAG_NVAL + q] = a;
    }
}

__global__ void k_diag_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    int q = threadIdx.x;
    if (q >= DIAG_NVAL) return;
    double a = partial[q];
    for (int b = 1; b < nblocks; b++) {
        double o = partial[(size_t)b * DIAG_NVAL + q];
        a = (q == 10) ? fmax(a, o) : a + o;
    }
    out[q] = a;
}

// weighted histogram of kinetic energy (quantity 0) or cos(theta_z) (quantity 1) over actives
template <int SP>
__global__ void k_histogram(PopView Q, long long n, int quantity, double lo, double hi, int nbins, int logscale, double* __restrict__ out) {
    extern __shared__ double bins[];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) bins[b] = 0;
    __syncthreads();
    double a = logscale ? log10(lo) : lo, bb = logscale ? log10(hi) : hi;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (!Q.active[i]) continue;
        Vec3 p = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
        double q = quantity == 0 ? kinenergy<SP>(p) : p.z / sqrt(dot(p, p));
        if (logscale) { if (!(q > 0)) continue; q = log10(q); }
        double f = (q - a) / (bb - a) * nbins;
        if (!(f >= 0) || !(f < nbins)) continue;
        atomicAdd(&bins[(int)f], Q.col[COL_W][i]);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) if (bins[b] != 0) atomicAdd(&out[b], bins[b]);
}

// ---- K4: roulette! (population.jl:291-309) and split! (:316-335) ------------------------------------------------------
// f(energy) of roulette!(f, popl) / split!(f, popl): tabulated by the host on n nodes uniform in E or log10(E) between lo
// and hi, interpolated linearly, flat outside; n == 1 is the constant law (the value travels in `c`).
struct EnergyLaw {
    double lo, hi, c;
    int n, logscale;
    const double* v;
};
__device__ __forceinline__ double law_value(const EnergyLaw& L, double eng) {
    if (L.n <= 1) return L.c;
    double x = L.logscale ? log10(eng) : eng;
    double u = (x - L.lo) / (L.hi - L.lo) * (double)(L.n - 1);
    if (!(u > 0)) return L.v[0];
    if (u >= (double)(L.n - 1)) return L.v[L.n - 1];
    int k = (int)u;
    double f = u - (double)k;
    return L.v[k] * (1 - f) + L.v[k + 1] * f;
}

static __global__ void k_roulette(PopView Q, long long n, EnergyLaw L, uint32_t step, uint32_t seed_lo, uint32_t seed_hi) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !Q.active[i]) return;
    Vec3 pp = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
    const double prob = law_value(L, kinenergy_rt(Q.species, pp));
    Rng rng;
    rng.init(Q.uid[i], DOM_ROULETTE);
    if (rng.u(step, seed_lo, seed_hi) < prob) Q.col[COL_W][i] /= prob;
    else Q.active[i] = 0;
}

static __global__ void k_split(PopView Q, long long n, EnergyLaw L, uint32_t step, uint32_t seed_lo, uint32_t seed_hi, int* flags) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int k = 0;
    bool act = i < n && Q.active[i];
    double w = 0;
    if (act) {
        Vec3 pp = {Q.col[COL_P0][i], Q.col[COL_P1][i], Q.col[COL_P2][i]};
        const double eng = kinenergy_rt(Q.species, pp);
        const double pmean = law_value(L, eng);
        w = Q.col[COL_W][i] / (1 + pmean);
        Q.col[COL_W][i] = w;
        Rng rng;
        rng.init(Q.uid[i], DOM_SPLIT);
        double u = rng.u(step, seed_lo, seed_hi);
        double pk = exp(-pmean), cdf = pk;   // Poisson(p) by sequential inversion
        while (u > cdf && k < 1000) { k++; pk *= pmean / k; cdf += pk; }
        if (!(eng > Q.energy_cut)) k = 0;   // add_particle! refuses copies at or below the cut (population.jl:105)
    }
    // warp-aggregated append of all copies
    int lane = threadIdx.x & 31;
    int incl = k;
    for (int off = 1; off < 32; off <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    int tot = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && tot) base = atomicAdd(Q.n, (unsigned long long)tot);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!act || k == 0) return;
    uint64_t uid = Q.uid[i];
    for (int c = 0; c < k; c++) {
        long long slot = (long long)base + (incl - k) + c;
        if (slot >= Q.capacity) { atomicOr(flags, PTL_ERR_CAPACITY_OVERFLOW); break; }
        uint64_t cu[2];
        child_uids(uid ^ ((uint64_t)DOM_SPLIT << 32), (uint32_t)c, step, seed_lo, seed_hi, cu);
#pragma unroll
        for (int q = 0; q < 10; q++) Q.col[q][slot] = Q.col[q][i];
        Q.active[slot] = 1;
        Q.uid[slot] = cu[0];
    }
}

// ---- shuffle! (population.jl:266-271): sort the rows by a 64-bit key drawn from the counter-based RNG -----------------
static __global__ void k_shuffle_keys(unsigned long long* __restrict__ keys, long long* __restrict__ rowid, long long n, uint32_t step,
                                      uint32_t seed_lo, uint32_t seed_hi) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(0u, step, seed_lo, seed_hi, (uint32_t)i, (uint32_t)((unsigned long long)i >> 32) ^ DOM_SHUFFLE, o);
    keys[i] = ((unsigned long long)o[1] << 32) | o[0];
    rowid[i] = i;
}
template <typename T>
static __global__ void k_gather(const T* __restrict__ src, const long long* __restrict__ idx, T* __restrict__ dst, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

}  // namespace ptl
