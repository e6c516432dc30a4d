// ptl_advance_wq.cuh — K1, warp-private variant of the wavefront advance kernel (the default for leptons in round 2).
//
// k_advance_bq (round 1) shares one pool of 512 slots between the 8 warps of a CTA: per-class lists in shared memory,
// appended with `__match_any_sync` + shared atomics, a chunk plan recomputed by every warp, and ONE CTA-wide barrier
// per round.  ncu on the driver bench (profiles/r1_s3_bench_main_kernel_ncu_summary.csv, r1_s3_bq_hotspots.md): 24 % of
// the warp time is spent at that barrier (2.3 stalled warps per issue), another 21 % of the warp samples and 23 % of the
// executed instructions are the scheduler itself, issue slots 41 % busy.
//
// Here every WARP owns WQ_K x 32 slots of the shared-memory pool and nobody else ever touches them:
//   * no CTA barrier after start-up, no shared atomics, no lists: the class of a slot is its state word;
//   * a round of a warp = every lane reads the state words of its WQ_K slots, the class populations are summed with two
//     REDUX adds on byte-packed counters, the warp picks ONE class (most pending entries, plus an age bonus so that rare
//     classes — finished particles waiting for write-back, bremsstrahlung — cannot starve), compacts up to 32 slot ids
//     of that class through 64 bytes of shared memory (ballot + popc prefix), and executes that unit with all lanes
//     coherent.  ~45 scheduler instructions per round instead of ~250;
//   * with all WQ_K x 32 slots always pending in one of three busy classes (STEP, COULOMB, RBEB) the fullest class
//     holds >= 1/3 of them and, because the classes that are not picked only grow, chunks run at 28-32 lanes.
// Work units, arithmetic, draw order and Philox streams are those of k_advance_wf (shared wf_execute_unit), so results
// are identical particle by particle to the other variants (tests/test_gpu_parity.py runs every kernel variant).
#pragma once
#include "ptl_advance_wf.cuh"

namespace ptl {

#ifndef WQ_THREADS_MACRO
#define WQ_THREADS_MACRO 256
#endif
#ifndef WQ_K_MACRO
#define WQ_K_MACRO 2                    // slots per lane
#endif
#ifndef WQ_MIN_BLOCKS
#define WQ_MIN_BLOCKS 2
#endif
#ifndef WQ_AGE_SHIFT
#define WQ_AGE_SHIFT 1                  // a class that was passed over gains (rounds waited << WQ_AGE_SHIFT) / 8 lanes of priority
#endif
constexpr int WQ_THREADS = WQ_THREADS_MACRO;
constexpr int WQ_WARPS = WQ_THREADS / 32;
constexpr int WQ_K = WQ_K_MACRO;
constexpr int WQ_NS = 32 * WQ_K;                    // slots per warp
constexpr int WQ_SLOTS = WQ_NS * WQ_WARPS;          // slots per CTA
static_assert(WQ_NS <= 255, "class populations are summed in 8-bit fields");

constexpr size_t WQ_POOL_BYTES =
    ((sizeof(double) * WD_NCOL * WQ_SLOTS + 16 * WQ_SLOTS + 4 * 5 * WQ_SLOTS + 2 * 32 * WQ_WARPS) + 15) / 16 * 16;

template <int SP, int TK, bool FIRST, bool CB>
__global__ void __launch_bounds__(WQ_THREADS, WQ_MIN_BLOCKS) k_advance_wq(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                              unsigned long long* row_counter, const long long* __restrict__ rows,
                                                              const unsigned long long* __restrict__ nrows) {
    if (rows != nullptr) { i0 = 0; i1 = (long long)*nrows; }     // index-list mode (rows deferred by the streaming kernel)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    WfPool S;
    S.np = WQ_SLOTS;
    unsigned char* ptr = smem_raw;
    S.d = reinterpret_cast<double*>(ptr); ptr += sizeof(double) * WD_NCOL * WQ_SLOTS;
    S.uid = reinterpret_cast<unsigned long long*>(ptr); ptr += 8 * WQ_SLOTS;
    S.row = reinterpret_cast<long long*>(ptr); ptr += 8 * WQ_SLOTS;
    S.idx = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WQ_SLOTS;
    S.cblock = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WQ_SLOTS;
    S.c2 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WQ_SLOTS;
    S.c3 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WQ_SLOTS;
    S.state = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * WQ_SLOTS;
    S.cnt = nullptr;
    S.order = nullptr;
    unsigned short* order_all = reinterpret_cast<unsigned short*>(ptr);                 // [warp][32]
    double* tsm = reinterpret_cast<double*>(smem_raw + WQ_POOL_BYTES);

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
    TS.rate = T.rate;
    TS.ratebound = tsm + nrate;
    TS.procs = reinterpret_cast<const ptl_process_desc*>(tsm + nrate + nrb);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int base = wid * WQ_NS;
    unsigned short* order = order_all + 32 * wid;
#pragma unroll
    for (int k = 0; k < WQ_K; k++) S.state[base + 32 * k + lane] = WS_LOAD;        // every slot starts as an empty LOAD item
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;
    __syncthreads();                                                                // the table is staged; last CTA-wide barrier

    // rounds a non-empty class has been passed over, 6 x 5 bits (saturating at 31)
    unsigned age = 0;
    unsigned round = 0;
    for (;; round++) {
        uint32_t cls[WQ_K];
        uint32_t pa = 0, pb = 0;
#pragma unroll
        for (int k = 0; k < WQ_K; k++) {
            cls[k] = S.state[base + 32 * k + lane] & 0xffu;
            pa += cls[k] < 4u ? (1u << (8 * cls[k])) : 0u;                          // LOAD, STEP, COULOMB, RBEB
            pb += (cls[k] == WS_IONFIN) ? 1u : (cls[k] == WS_OTHER ? 256u : 0u);
        }
        pa = __reduce_add_sync(0xffffffffu, pa);
        pb = __reduce_add_sync(0xffffffffu, pb);
        if ((pa | pb) == 0u) break;                                                 // every slot of this warp is retired
        // pick the class: most pending entries (capped at a full chunk) plus the age bonus; ties -> lower class
        int best = 0, bscore = -1;
#pragma unroll
        for (int c = 0; c < WS_IDLE; c++) {
            const int n = (int)(((c < 4 ? pa : pb) >> (8 * (c & 3))) & 0xffu);
            const int a = (int)((age >> (5 * c)) & 31u);
            const int score = n > 0 ? ((n < 32 ? n : 32) << 3) + (a << WQ_AGE_SHIFT) : -1;
            if (score > bscore) { bscore = score; best = c; }
        }
        // ages: the picked class restarts, every other non-empty class waits one more round
        {
            unsigned na = 0;
#pragma unroll
            for (int c = 0; c < WS_IDLE; c++) {
                const unsigned n = ((c < 4 ? pa : pb) >> (8 * (c & 3))) & 0xffu;
                unsigned a = (age >> (5 * c)) & 31u;
                a = (c == best || n == 0u) ? 0u : (a < 31u ? a + 1u : 31u);
                na |= a << (5 * c);
            }
            age = na;
        }
        // compact up to 32 slot ids of the picked class; the starting group rotates so that no slot waits forever when
        // its class stays above 32 entries
        int pos = 0;
#pragma unroll
        for (int kk = 0; kk < WQ_K; kk++) {
            const int k = (WQ_K == 1) ? 0 : (int)((kk + round) % WQ_K);
            uint32_t ck = cls[0];
#pragma unroll
            for (int q = 1; q < WQ_K; q++) ck = (q == k) ? cls[q] : ck;
            const bool mine = ck == (uint32_t)best;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            const int q = pos + __popc(m & ltmask);
            if (mine && q < 32) order[q] = (unsigned short)(base + 32 * k + lane);
            pos += __popc(m);
        }
        __syncwarp();
        const int cnt = pos < 32 ? pos : 32;
        const bool has = lane < cnt;
        const unsigned amask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
        if (has) {
            const int it = (int)order[lane];
            const uint32_t sw = S.state[it];
            wf_execute_unit<SP, TK, FIRST, CB>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, amask, lane, ltmask, row_counter, i0, i1, nsub, rows);
        }
        __syncwarp();
    }

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
    if (lane == 0) { atomicMax(P.dbg, (unsigned long long)round); atomicAdd(P.dbg + 1, (unsigned long long)round); atomicAdd(P.dbg + 2, 1ULL); }
}

}  // namespace ptl
