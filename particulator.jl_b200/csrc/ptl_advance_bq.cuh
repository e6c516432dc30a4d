// ptl_advance_bq.cuh — K1, list-scheduled variant of the wavefront advance kernel (the default for leptons).
//
// k_advance_wf re-sorts every slot of the CTA every round (ballots, count matrix, two extra barriers: ~250 scheduler
// instructions per warp and round = 30 % of all instructions in ncu, profiles/r1_v25_*).  Here the sort is incremental:
// when a lane finishes a work unit it appends its slot to the shared-memory list of the slot's NEXT class (one
// atomicAdd per destination class and warp, `__match_any_sync`), so at the single barrier that ends a round the
// per-class lists of the next round are already built.  A round is then: read six counters, pick this warp's chunks
// (32 consecutive entries of one class; full chunks first, then the largest remainders; what is not picked is carried
// over and only grows), execute, append, barrier.  The pool holds BQ_SLOTS = 2 x threads slots and every warp
// executes up to two chunks per round, which halves the barriers per unit and keeps the chunks full.
// Work units, arithmetic, draw order and Philox streams are those of k_advance_wf (shared wf_execute_unit).
#pragma once
#include "ptl_advance_wf.cuh"

namespace ptl {

#ifndef BQ_THREADS_MACRO
#define BQ_THREADS_MACRO 256
#endif
constexpr int BQ_THREADS = BQ_THREADS_MACRO;
constexpr int BQ_WARPS = BQ_THREADS / 32;
#ifndef BQ_CPW_MACRO
#define BQ_CPW_MACRO 2
#endif
#ifndef BQ_MIN_BLOCKS
#define BQ_MIN_BLOCKS 2
#endif
#ifndef BQ_PART_MIN
#define BQ_PART_MIN 28
#endif
#ifndef BQ_QUIET_DIV
#define BQ_QUIET_DIV 2                  // the pool counts as quiet when fewer than 1/BQ_QUIET_DIV of the chunk slots are full chunks
#endif
constexpr int BQ_CPW = BQ_CPW_MACRO;               // chunks per warp and round
#ifdef BQ_SLOTS_MACRO
constexpr int BQ_SLOTS = BQ_SLOTS_MACRO;            // may be below the chunk capacity (then not every chunk slot is used)
#else
constexpr int BQ_SLOTS = BQ_THREADS * BQ_CPW;
#endif
static_assert(BQ_SLOTS <= BQ_THREADS * BQ_CPW && BQ_SLOTS % 32 == 0, "the plan executes at most BQ_CHUNKS full chunks per round");
constexpr int BQ_LIST_BYTES = 2 * 2 * WS_IDLE * BQ_SLOTS;
// (Measured and dropped: class lists as circular queues -- consumed at the head, appended at the tail, no carry-over copies,
// pools of 512..608 slots.  14 % slower at 512 slots and it needs rings of twice the pool: a slot that returns to its own
// class (null collision, rejected RBEB trial) is appended while its old entry is still being read.)
constexpr int BQ_NCLASS = WS_IDLE;                 // 6 lists
constexpr int BQ_CHUNKS = BQ_WARPS * BQ_CPW;       // chunks executed per round

constexpr size_t BQ_POOL_BYTES =
    ((sizeof(double) * WD_NCOL * BQ_SLOTS + 16 * BQ_SLOTS + 4 * 5 * BQ_SLOTS + BQ_LIST_BYTES + 4 * 3 * 8) + 15) / 16 * 16;

template <int SP, int TK, bool FIRST, bool CB>
__global__ void __launch_bounds__(BQ_THREADS, BQ_MIN_BLOCKS) k_advance_bq(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                              unsigned long long* row_counter, const long long* __restrict__ rows,
                                                              const unsigned long long* __restrict__ nrows) {
    if (rows != nullptr) { i0 = 0; i1 = (long long)*nrows; }     // index-list mode (rows deferred by the streaming kernel)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    WfPool S;
    S.np = BQ_SLOTS;
    unsigned char* ptr = smem_raw;
    S.d = reinterpret_cast<double*>(ptr); ptr += sizeof(double) * WD_NCOL * BQ_SLOTS;
    S.uid = reinterpret_cast<unsigned long long*>(ptr); ptr += 8 * BQ_SLOTS;
    S.row = reinterpret_cast<long long*>(ptr); ptr += 8 * BQ_SLOTS;
    S.idx = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * BQ_SLOTS;
    S.cblock = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * BQ_SLOTS;
    S.c2 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * BQ_SLOTS;
    S.c3 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * BQ_SLOTS;
    S.state = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * BQ_SLOTS;
    S.cnt = nullptr;
    S.order = nullptr;
    unsigned short* lists = reinterpret_cast<unsigned short*>(ptr); ptr += BQ_LIST_BYTES;   // [2][class][slot]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(ptr);                                               // [3][8]
    double* tsm = reinterpret_cast<double*>(smem_raw + BQ_POOL_BYTES);

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
    // round 0: every slot is an empty LOAD item
    for (int q = tid; q < BQ_SLOTS; q += blockDim.x) { lists[q] = (unsigned short)q; S.state[q] = WS_LOAD; }
    if (tid < 24) cnt[tid] = (tid == WS_LOAD) ? BQ_SLOTS : 0;
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;
    __syncthreads();

    unsigned round = 0;
    for (;; round++) {
        const unsigned cb3 = round % 3;
        const unsigned int* ccur = cnt + 8 * cb3;
        unsigned int* cnxt = cnt + 8 * ((cb3 + 1) % 3);
        unsigned int* cold = cnt + 8 * ((cb3 + 2) % 3);          // read in the previous round: safe to clear now
        const unsigned short* lcur = lists + (round & 1) * (BQ_NCLASS * BQ_SLOTS);
        unsigned short* lnxt = lists + ((round & 1) ^ 1) * (BQ_NCLASS * BQ_SLOTS);
        if (tid < 8) cold[tid] = 0;
#ifdef BQ_PROFILE
        const long long pr_t0 = clock64();
        long long pr_umax = 0;
#endif

        // ---- chunk plan, computed once per warp with lane c holding class c ----
        // full chunks in class order, then the largest remainders (ties -> lower class) while chunks are left
        const int n_c = lane < BQ_NCLASS ? (int)ccur[lane] : 0;
        const int f_c = n_c >> 5, r_c = n_c & 31;
        int incl = f_c;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const int start_c = incl - f_c;                          // first plan index of class c's full chunks
        const int nfull = __shfl_sync(0xffffffffu, incl, BQ_NCLASS - 1);
        int tot = n_c;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
        if (__shfl_sync(0xffffffffu, tot, 0) == 0) break;        // block-uniform: every slot retired
#ifdef PTL_DEBUG_TAIL
        if (wid == 0 && __shfl_sync(0xffffffffu, tot, 0) <= 2) {
            if (lane < BQ_NCLASS && n_c > 0) {
                atomicAdd(P.dbg + 4 + lane, 1ULL);
                if ((round & 1023) == 0) {
                    const int it0 = lcur[lane * BQ_SLOTS];
                    P.dbg[3] = (unsigned long long)S.row[it0];
                    P.dbg[10] = __double_as_longlong(WFD(WD_P0, it0)); P.dbg[11] = __double_as_longlong(WFD(WD_P0 + 1, it0));
                    P.dbg[12] = __double_as_longlong(WFD(WD_P0 + 2, it0)); P.dbg[13] = __double_as_longlong(WFD(WD_R, it0));
                    P.dbg[14] = __double_as_longlong(WFD(WD_S, it0));
                }
            }
        }
#endif
        int rk_c = 0;
#pragma unroll
        for (int d = 0; d < BQ_NCLASS; d++) {
            const int rd = __shfl_sync(0xffffffffu, r_c, d);
            rk_c += (rd > r_c) || (rd == r_c && d < lane);
        }
        // (A pool larger than one round executes -- 544..608 slots for 16 chunk slots, so that every round finds 16 full
        // chunks -- was measured 3-6 % slower: the carry-over copies grow with the waiting entries.)
        const int npartial = BQ_CHUNKS - nfull;                  // nfull <= BQ_CHUNKS because the pool has BQ_CHUNKS*32 slots
        // While the pool is busy (at least half of the chunk slots are full chunks) a remainder below BQ_PART_MIN lanes waits
        // and grows instead of costing a whole warp pass for a few lanes (+0.7 %; RBEB/IONFIN chunks ran at 14-17 lanes).
        // With a quiet pool every remainder runs, so nothing can starve at the tail.
        const bool part_c = lane < BQ_NCLASS && r_c > 0 && rk_c < npartial && (r_c >= BQ_PART_MIN || BQ_QUIET_DIV * nfull < BQ_CHUNKS);
        const int done_c = f_c * 32 + (part_c ? r_c : 0);        // entries of class c executed this round

        // carry-over: what this round does not execute moves to the next round's lists (warp c handles class c)
        for (int c = wid; c < BQ_NCLASS; c += BQ_WARPS) {
            const int nn = __shfl_sync(0xffffffffu, n_c, c), done = __shfl_sync(0xffffffffu, done_c, c);
            const int left = nn - done;
            if (left > 0) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&cnxt[c], (unsigned)left);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int q = lane; q < left; q += 32) lnxt[c * BQ_SLOTS + base + q] = lcur[c * BQ_SLOTS + done + q];
            }
        }

        // ---- execute this warp's chunks ----
        // Static assignment (chunk = wid + cw * BQ_WARPS).  Measured alternative, -DBQ_DYNAMIC_CHUNKS: warps claim plan
        // indices from a shared counter so that a warp which drew short units takes the next chunk instead of waiting at
        // the barrier (the wait is 21 % of the warp samples, profiles/r1_s2_bq_hotspots.md).  On B200 it is 2 % SLOWER
        // (36.7 vs 35.9 ms, 4e6 electrons): the second CTA of the SM already fills the barrier wait, and the claim adds
        // an atomic + shuffle per chunk.  Kept as a documented negative result.  Also measured: two consecutive plan
        // entries per warp (chunk = wid * BQ_CPW + cw; mostly one class per warp, warmer instruction cache) 17 % slower
        // (41.4 ms: a warp with two STEP chunks holds the barrier); pairing first with last (w, 15 - w) +0.5 %, within noise.
#ifndef BQ_DYNAMIC_CHUNKS
#pragma unroll 1
        for (int cw = 0; cw < BQ_CPW; cw++) {
            const int chunk = wid + cw * BQ_WARPS;               // index in the plan order
#else
        const int nplan = nfull + __popc(__ballot_sync(0xffffffffu, part_c));
        unsigned int* claim = cnt + 8 * cb3 + 6;                 // unused entry of this round's counter row (zeroed two rounds ago)
#pragma unroll 1
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = (int)atomicAdd(claim, 1u);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= nplan) break;
#endif
            const unsigned mfull = __ballot_sync(0xffffffffu, lane < BQ_NCLASS && chunk >= start_c && chunk < start_c + f_c);
            const unsigned mpart = __ballot_sync(0xffffffffu, part_c && rk_c == chunk - nfull);
            const unsigned msel = mfull ? mfull : mpart;
            int it = -1;
            if (msel) {
                const int my_c = __ffs(msel) - 1;
                const int k0 = mfull ? (chunk - __shfl_sync(0xffffffffu, start_c, my_c)) : __shfl_sync(0xffffffffu, f_c, my_c);
                const int nn = __shfl_sync(0xffffffffu, n_c, my_c);
                const int pos = k0 * 32 + lane;
                if (pos < nn) it = (int)lcur[my_c * BQ_SLOTS + pos];
            }
            const bool has = it >= 0;
            const unsigned amask = __ballot_sync(0xffffffffu, has);
#ifdef BQ_PROFILE
            const long long pr_u0 = clock64();
#endif
            if (has) {
                const uint32_t sw = S.state[it];
                wf_execute_unit<SP, TK, FIRST, CB>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, amask, lane, ltmask, row_counter, i0, i1, nsub, rows);
                __syncwarp(amask);
                const int nc = (int)(S.state[it] & 0xffu);       // next class; IDLE slots are retired
                const unsigned grp = __match_any_sync(amask, nc);
                if (nc < BQ_NCLASS) {
                    const int leader = __ffs(grp) - 1;
                    unsigned pos = 0;
                    if (lane == leader) pos = atomicAdd(&cnxt[nc], (unsigned)__popc(grp));
                    pos = __shfl_sync(grp, pos, leader) + __popc(grp & ltmask);
                    lnxt[nc * BQ_SLOTS + pos] = (unsigned short)it;
                }
            }
            __syncwarp();
#ifdef BQ_PROFILE
            if (msel) {
                const long long du = clock64() - pr_u0;
                pr_umax = du > pr_umax ? du : pr_umax;
                if (lane == 0) {
                    const int pc = __ffs(msel) - 1;
                    atomicAdd(P.dbg + 16 + pc, (unsigned long long)du); atomicAdd(P.dbg + 24 + pc, 1ULL); atomicMax(P.dbg + 32 + pc, (unsigned long long)du);
                }
            }
#endif
        }
#ifdef BQ_PROFILE
        const long long pr_b0 = clock64();
#endif
        __syncthreads();
#ifdef BQ_PROFILE
        if (lane == 0) {
            const long long pr_t1 = clock64();
            atomicAdd(P.dbg + 40, (unsigned long long)(pr_t1 - pr_b0)); atomicAdd(P.dbg + 41, 1ULL);
            atomicAdd(P.dbg + 42, (unsigned long long)(pr_t1 - pr_t0)); atomicAdd(P.dbg + 43, (unsigned long long)pr_umax);
        }
#endif
    }

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
    if (tid == 0) { atomicMax(P.dbg, (unsigned long long)round); atomicAdd(P.dbg + 1, (unsigned long long)round); atomicAdd(P.dbg + 2, 1ULL); }
}

}  // namespace ptl
