// ptl_advance_aq.cuh — K1, queue-driven variant of the wavefront advance kernel.
//
// k_advance_wf runs lock-step rounds: sort all slots of the CTA, execute one chunk per warp, barrier.  ncu showed
// the cost of that structure (profiles/r1_v23_*): ~250 scheduler instructions per warp and round, and warps that
// executed a short unit waiting at the barrier for the warp with the longest one (CTA-barrier stalls = 3.4 warps per
// issue slot, issue slots 35 % busy).  Here the warps of a CTA are autonomous:
//   * the CTA still owns a shared-memory pool of particle slots (AQ_SLOTS > threads, so queues stay deep);
//   * every work class has a ring buffer of slot ids in shared memory with monotonic head / tail counters;
//   * a warp claims up to 32 ids of ONE class (compare-and-swap on the head), executes that unit for them with all
//     lanes coherent, and pushes every slot onto the ring of its next class (one atomicAdd per destination class and
//     warp, `__match_any_sync`); no CTA-wide barrier exists after start-up;
//   * a warp exits when every ring is empty and no other warp holds claimed items.
// Rings carry slot+1 (0 = empty cell): a consumer that claimed a position spins until the producer's store lands, a
// producer waits for a wrapped cell to be cleared; both waits are bounded by the other side's straight-line code.  A
// watchdog turns a (never observed) stuck warp into a sticky error flag instead of a hang.
// Work units, arithmetic, draw order and Philox streams are those of k_advance_wf (shared wf_execute_unit).
#pragma once
#include "ptl_advance_wf.cuh"

namespace ptl {

constexpr int AQ_THREADS = 256;
constexpr int AQ_WARPS = AQ_THREADS / 32;
constexpr int AQ_SLOTS = 512;
constexpr int AQ_NCLASS = WS_IDLE;   // 6 rings

constexpr size_t AQ_POOL_BYTES =
    ((sizeof(double) * WD_NCOL * AQ_SLOTS + 16 * AQ_SLOTS + 4 * 5 * AQ_SLOTS + 2 * AQ_NCLASS * AQ_SLOTS + 4 * 32) + 15) / 16 * 16;

struct AqQueues {
    unsigned short* ring;          // [AQ_NCLASS][AQ_SLOTS], slot + 1
    volatile unsigned int* head;   // [AQ_NCLASS] claimed
    volatile unsigned int* tail;   // [AQ_NCLASS] reserved by producers
    volatile unsigned int* inflight;
};

template <int SP, int TK, bool FIRST, bool CB>
__global__ void __launch_bounds__(AQ_THREADS, 2) k_advance_aq(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                              unsigned long long* row_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    WfPool S;
    S.np = AQ_SLOTS;
    unsigned char* ptr = smem_raw;
    S.d = reinterpret_cast<double*>(ptr); ptr += sizeof(double) * WD_NCOL * AQ_SLOTS;
    S.uid = reinterpret_cast<unsigned long long*>(ptr); ptr += 8 * AQ_SLOTS;
    S.row = reinterpret_cast<long long*>(ptr); ptr += 8 * AQ_SLOTS;
    S.idx = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * AQ_SLOTS;
    S.cblock = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * AQ_SLOTS;
    S.c2 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * AQ_SLOTS;
    S.c3 = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * AQ_SLOTS;
    S.state = reinterpret_cast<uint32_t*>(ptr); ptr += 4 * AQ_SLOTS;
    S.cnt = nullptr;
    S.order = nullptr;
    AqQueues A;
    A.ring = reinterpret_cast<unsigned short*>(ptr); ptr += 2 * AQ_NCLASS * AQ_SLOTS;
    unsigned int* ctr = reinterpret_cast<unsigned int*>(ptr);
    A.head = ctr; A.tail = ctr + 8; A.inflight = ctr + 16;
    double* tsm = reinterpret_cast<double*>(smem_raw + AQ_POOL_BYTES);

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

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    // every slot starts as an empty LOAD item
    for (int q = tid; q < AQ_NCLASS * AQ_SLOTS; q += blockDim.x) A.ring[q] = (q < AQ_SLOTS) ? (unsigned short)(q + 1) : (unsigned short)0;
    for (int q = tid; q < AQ_SLOTS; q += blockDim.x) S.state[q] = WS_LOAD;
    if (tid < 32) ctr[tid] = (tid == 8 + WS_LOAD) ? AQ_SLOTS : 0;      // tail[LOAD] = AQ_SLOTS
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;
    __syncthreads();                     // the only CTA-wide barrier

    unsigned idle_polls = 0;
    for (;;) {
        // ---- claim up to 32 items of one class ----
        // lanes 0..5 read the depth of "their" ring in parallel; the deepest ring wins.  A warp prefers to wait a few
        // hundred ns for a full chunk while other warps still hold items (they are about to push), and only takes a
        // partial chunk when nobody is processing or the wait budget is spent: keeps ~30 lanes busy per instruction.
        int c_sel = -1;
        unsigned h_sel = 0, k_sel = 0;
        for (int attempt = 0; attempt < 24; attempt++) {
            unsigned av = 0;
            if (lane < AQ_NCLASS) av = A.tail[lane] - A.head[lane];
            unsigned key = (av << 3) | (unsigned)(7 - lane);          // deepest ring, ties -> lowest class
            unsigned best = key;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) { unsigned o = __shfl_xor_sync(0xffffffffu, best, off); best = o > best ? o : best; }
            best = __shfl_sync(0xffffffffu, best, 0);                  // lanes 0..7 hold the maximum of their group of 8
            const unsigned bestn = best >> 3;
            const int bc = 7 - (int)(best & 7u);
            if (bestn == 0) break;
            const unsigned infl = *A.inflight;
            if (bestn < 32 && infl > 0 && attempt < 23) { __nanosleep(40); continue; }
            int ok = 0;
            unsigned old = 0, kk = 0;
            if (lane == 0) {
                unsigned k = bestn < 32 ? bestn : 32;
                atomicAdd((unsigned int*)A.inflight, k);
                old = A.head[bc];
                unsigned av2 = A.tail[bc] - old;
                kk = av2 < k ? av2 : k;
                if (kk > 0 && atomicCAS((unsigned int*)&A.head[bc], old, old + kk) == old) {
                    if (kk < k) atomicSub((unsigned int*)A.inflight, k - kk);
                    ok = 1;
                } else {
                    atomicSub((unsigned int*)A.inflight, k);
                }
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (ok) { c_sel = bc; h_sel = old; k_sel = kk; break; }
        }
        c_sel = __shfl_sync(0xffffffffu, c_sel, 0);
        if (c_sel < 0) {
            // nothing claimable: done when no ring holds items and nobody is processing
            unsigned busy = 0;
            if (lane == 0) {
                busy = *A.inflight;
#pragma unroll
                for (int c = 0; c < AQ_NCLASS; c++) busy += A.tail[c] - A.head[c];
            }
            busy = __shfl_sync(0xffffffffu, busy, 0);
            if (busy == 0) {                       // look twice: a claim in flight can hide items for an instant
                __nanosleep(256);
                if (lane == 0) {
                    busy = *A.inflight;
#pragma unroll
                    for (int c = 0; c < AQ_NCLASS; c++) busy += A.tail[c] - A.head[c];
                }
                busy = __shfl_sync(0xffffffffu, busy, 0);
                if (busy == 0) break;
            }
            if (++idle_polls > 20000000u) { if (lane == 0) atomicOr(P.flags, PTL_ERR_NAN_STATE); break; }   // watchdog
            __nanosleep(128);
            continue;
        }
        idle_polls = 0;
        h_sel = __shfl_sync(0xffffffffu, h_sel, 0);
        k_sel = __shfl_sync(0xffffffffu, k_sel, 0);
        const bool has = (unsigned)lane < k_sel;
        const unsigned amask = __ballot_sync(0xffffffffu, has);
        if (has) {
            volatile unsigned short* cell = A.ring + c_sel * AQ_SLOTS + ((h_sel + lane) & (AQ_SLOTS - 1));
            unsigned v;
            unsigned spins = 0;
            while ((v = *cell) == 0) { if (++spins > 100000000u) { atomicOr(P.flags, PTL_ERR_NAN_STATE); break; } }
            *cell = 0;
            __threadfence_block();
            const int it = (int)v - 1;
            int nc = AQ_NCLASS;
            if (it >= 0) {
                const uint32_t sw = S.state[it];
                wf_execute_unit<SP, TK, FIRST, CB>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, amask, lane, ltmask, row_counter, i0, i1, nsub);
                nc = (int)(S.state[it] & 0xffu);
            }
            __threadfence_block();
            __syncwarp(amask);
            // ---- push to the ring of the next class (IDLE slots are retired) ----
            const unsigned grp = __match_any_sync(amask, nc);
            if (nc < AQ_NCLASS) {
                const int leader = __ffs(grp) - 1;
                unsigned pos = 0;
                if (lane == leader) pos = atomicAdd((unsigned int*)&A.tail[nc], (unsigned)__popc(grp));
                pos = __shfl_sync(grp, pos, leader) + __popc(grp & ltmask);
                volatile unsigned short* dst = A.ring + nc * AQ_SLOTS + (pos & (AQ_SLOTS - 1));
                unsigned spins2 = 0;
                while (*dst != 0) { if (++spins2 > 100000000u) { atomicOr(P.flags, PTL_ERR_NAN_STATE); break; } }
                *dst = (unsigned short)(it + 1);
            }
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicSub((unsigned int*)A.inflight, k_sel); }
    }

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
}

}  // namespace ptl
