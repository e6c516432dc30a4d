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
//     coherent.  181 scheduler instructions per round in SASS (ncu source page; the first estimate here said ~45), 19 % of
//     what the kernel executes; the attempts to shrink them are below (WQ_FULL_RECOUNT=0);
//   * with all WQ_K x 32 slots always pending in one of three busy classes (STEP, COULOMB, RBEB) the fullest class
//     holds >= 1/3 of them and, because the classes that are not picked only grow, chunks run at 28-32 lanes.
// Work units, arithmetic, draw order and Philox streams are those of k_advance_wf (shared wf_execute_unit), so results
// are identical particle by particle to the other variants (tests/test_gpu_parity.py runs every kernel variant).
#pragma once
#include "ptl_advance_wf.cuh"

namespace ptl {

#ifndef WQ_THREADS_MACRO
#define WQ_THREADS_MACRO 512
#endif
#ifndef WQ_K_MACRO
#define WQ_K_MACRO 3                    // slots per lane
#endif
#ifndef WQ_MIN_BLOCKS
#define WQ_MIN_BLOCKS 1
#endif
#ifndef WQ_AGE_SHIFT
#define WQ_AGE_SHIFT 1                  // (autonomous mode) a class that was passed over gains (rounds waited << WQ_AGE_SHIFT) / 8 lanes of priority
#endif
#ifndef WQ_ALOAD
#define WQ_ALOAD 0                      // 1: asynchronous LOAD (cp.async the row into the slot, finish it in a later round, class LOADWAIT).
                                        // Measured on B200 (4e6 electrons): 29.9 ms against 28.2 ms for the plain LOAD — the extra, poorly filled
                                        // LOADWAIT rounds cost more than the long-scoreboard stall they remove (other warps already hide it).
#endif
#ifndef WQ_LOAD_WEIGHT
#define WQ_LOAD_WEIGHT 8                // score per pending LOAD / LOADWAIT entry (8 = like every other class)
#endif
#ifndef WQ_COUNT64
#define WQ_COUNT64 0                    // 1: one 64-bit one-hot add per slot instead of two 32-bit selects (measured: no faster)
#endif
#ifndef WQ_SYNC
#define WQ_SYNC 0                       // 0: autonomous warps (default); 1: the warps of a group agree on one class per round (one barrier)
#endif
#ifndef WQ_GROUP
#define WQ_GROUP (WQ_THREADS_MACRO / 32)   // (synchronous mode) warps that agree on one class per round; default: the whole CTA
#endif
#ifndef WQ_GROUP_STRIDED
#define WQ_GROUP_STRIDED 0              // 1: group = warps with the same (warp id mod number of groups), i.e. one SM sub-partition each
#endif
#ifndef WQ_CHUNKS
#define WQ_CHUNKS 1                     // chunks of 32 entries of the picked class a round may execute (lean scheduler only)
#endif
#ifndef WQ_CHUNK_MIN
#define WQ_CHUNK_MIN 24                 // a further chunk is taken only when it has at least this many entries
#endif
#ifndef WQ_FULL_RECOUNT
#define WQ_FULL_RECOUNT 1               // 1 (default): every lane re-counts its WQ_K slots every round; 0: the "lean" incremental scheduler below
                                        // (fewer instructions, measured 10 % SLOWER: 30.3 against 27.5 ms, see the comment at its loop)
#endif
#ifndef WQ_AGE_FORCE
#define WQ_AGE_FORCE 12                 // rounds after which a waiting class becomes the CTA's top class whatever its size
#endif
#ifndef WQ_FOLLOW_SLACK
#define WQ_FOLLOW_SLACK 12              // a warp follows the CTA's class unless its own best class has this many more lanes
#endif
constexpr int WQ_THREADS = WQ_THREADS_MACRO;
constexpr int WQ_WARPS = WQ_THREADS / 32;
constexpr int WQ_K = WQ_K_MACRO;
constexpr int WQ_NS = 32 * WQ_K;                    // slots per warp
constexpr int WQ_SLOTS = WQ_NS * WQ_WARPS;          // slots per CTA
static_assert(WQ_NS <= 255, "class populations are summed in 8-bit fields");

constexpr size_t WQ_POOL_BYTES =
    ((sizeof(double) * WD_NCOL * WQ_SLOTS + 16 * WQ_SLOTS + 4 * 5 * WQ_SLOTS + 2 * 32 * WQ_CHUNKS * WQ_WARPS + 8 * 4 * 8) + 15) / 16 * 16;

// carve the pool out of the dynamic shared memory (compile-time offsets: the kernel and the out-of-line OTHER unit agree)
__device__ __forceinline__ WfPool wq_pool(unsigned char* ptr) {
    WfPool S;
    S.np = WQ_SLOTS;
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
    return S;
}
// shared-memory table layout behind the pool: [cumulative rates | rate bound | process descriptors]
__device__ __forceinline__ void wq_table_layout(const TableView& T, const int TK, const bool fastsel, int& nrate, int& nrb) {
    nrate = (TK == 0) ? (fastsel ? WF_CUM_STRIDE * (T.k + 1) : T.order * T.nprocs * (T.k + 1)) : 0;
    nrb = (TK == 0) ? T.order * (T.k + 1) : 0;
}

#ifndef WQ_OTHER_OUTLINE
#define WQ_OTHER_OUTLINE 0              // 1: the OTHER unit is a real function (keeps ~2000 cold instructions out of the kernel body)
#endif
// FS: 1 = the table has the fast-selection shape (Chebyshev, order 3, <= 16 processes; checked by the launcher), so every
// `fastsel ? :` in the units folds at compile time; 0 = decided at run time from the table (any other table).
template <int SP, int TK, int FS>
static __device__ __noinline__ unsigned long long wq_other_unit(const AdvanceParams* Pp, int it, uint32_t sw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const AdvanceParams& P = *Pp;
    const TableView& T = P.tab[SP];
    const bool fastsel = FS ? true : ((TK == 0) && T.order == 3 && T.nprocs <= 16);
    int nrate, nrb;
    wq_table_layout(T, TK, fastsel, nrate, nrb);
    const WfPool S = wq_pool(smem_raw);
    double* tsm = reinterpret_cast<double*>(smem_raw + WQ_POOL_BYTES);
    SmemTable TS;
    TS.rate = T.rate;
    TS.ratebound = tsm + nrate;
    TS.procs = reinterpret_cast<const ptl_process_desc*>(tsm + nrate + nrb);
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    return wf_other_unit<SP, TK>(P, P.pop[SP], S, TS, fastsel, rc, P.pop[SP].energy_cut, it, sw);
}

#ifndef WQ_LOAD_OUTLINE
#define WQ_LOAD_OUTLINE 0               // 1: the LOAD unit is a real function too
#endif
template <int SP, int TK, bool FIRST, int FS>
static __device__ __noinline__ void wq_load_unit(const AdvanceParams* Pp, int it, uint32_t sw, unsigned ldmask, unsigned long long* row_counter,
                                                 long long i0, long long i1, const long long* rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const AdvanceParams& P = *Pp;
    const TableView& T = P.tab[SP];
    const bool fastsel = FS ? true : ((TK == 0) && T.order == 3 && T.nprocs <= 16);
    int nrate, nrb;
    wq_table_layout(T, TK, fastsel, nrate, nrb);
    const WfPool S = wq_pool(smem_raw);
    double* tsm = reinterpret_cast<double*>(smem_raw + WQ_POOL_BYTES);
    SmemTable TS;
    TS.rate = T.rate;
    TS.ratebound = tsm + nrate;
    TS.procs = reinterpret_cast<const ptl_process_desc*>(tsm + nrate + nrb);
    const int lane = threadIdx.x & 31;
    wf_load_unit<SP, TK, FIRST, false>(P, T, P.pop[SP], S, TS, fastsel, P.pop[SP].energy_cut, it, sw, ldmask, lane, (1u << lane) - 1u, row_counter, i0, i1, rows);
}

template <int SP, int TK, bool FIRST, bool CB, int FS = 0>
__global__ void __launch_bounds__(WQ_THREADS, WQ_MIN_BLOCKS) k_advance_wq(const __grid_constant__ AdvanceParams P, long long i0, long long i1,
                                                              unsigned long long* row_counter, const long long* __restrict__ rows,
                                                              const unsigned long long* __restrict__ nrows) {
    if (rows != nullptr) { i0 = 0; i1 = (long long)*nrows; }     // index-list mode (rows deferred by the streaming kernel)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TableView& T = P.tab[SP];
    const PopView& Q = P.pop[SP];
    const WfPool S = wq_pool(smem_raw);
    unsigned char* ptr = smem_raw + sizeof(double) * WD_NCOL * WQ_SLOTS + 16 * WQ_SLOTS + 4 * 5 * WQ_SLOTS;
    unsigned short* order_all = reinterpret_cast<unsigned short*>(ptr);                 // [warp][32]
    double* tsm = reinterpret_cast<double*>(smem_raw + WQ_POOL_BYTES);

    const bool fastsel = FS ? true : ((TK == 0) && T.order == 3 && T.nprocs <= 16);
    int nrate, nrb;
    wq_table_layout(T, TK, fastsel, nrate, nrb);
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
    unsigned short* order = order_all + 32 * WQ_CHUNKS * wid;
#pragma unroll
    for (int k = 0; k < WQ_K; k++) S.state[base + 32 * k + lane] = WS_LOAD;        // every slot starts as an empty LOAD item
    const RngCtx rc = {P.step, P.seed_lo, P.seed_hi};
    const double cut = Q.energy_cut;
    unsigned long long nsub = 0;
    __syncthreads();                                                                // the table is staged; last CTA-wide barrier

    // Scheduler state is lane-parallel: lane c (c < 6) owns class c — its population, its score, its age (rounds a
    // non-empty class has been passed over).  Decisions are one REDUX each (the first version evaluated every class in every
    // lane: 22 % of all instructions).
    int age = 0;
    unsigned round = 0;
    const int myc = lane < WS_IDLE ? lane : 0;
#if WQ_SYNC
    // CTA-synchronous class choice.  With autonomous warps the 16 warps of an SM execute up to five different work units
    // at the same time: the hot SASS (~30 KB for 99 % of the executed instructions, scattered over 158 KB) against a
    // 32 KB instruction cache, and ncu showed the barrier stall of the list-scheduled kernel (2.3 warps per issue) simply
    // replaced by instruction-fetch stalls (3.3; hit rate 73 %, profiles/r2_wq_async_ncu_summary.csv).  So the warps of a
    // CTA agree on ONE class per round: every warp adds its (capped) class populations to a shared 64-bit counter, one
    // barrier, everybody derives the same "top" class from the same totals, and a warp follows it unless its own pool
    // is clearly better served by another class.  All warps then run the same straight-line unit at the same time
    // (equal durations: the barrier wait is short, and the SM holds one or two units' code instead of five).
    // The agreement group is WQ_GROUP warps: the whole CTA, or (WQ_GROUP_STRIDED) the warps that share one of the four
    // SM sub-partitions (warp id mod 4 picks the scheduler), synchronised with a named barrier of their own.
    constexpr int NGROUPS = WQ_WARPS / WQ_GROUP;
    const int grp = WQ_GROUP_STRIDED ? (wid % NGROUPS) : (wid / WQ_GROUP);
    unsigned long long* csum = reinterpret_cast<unsigned long long*>(order_all + 32 * WQ_CHUNKS * WQ_WARPS) + 4 * grp;     // [group][3(+1)], 6 x 10-bit fields
    if (tid < 4 * NGROUPS) reinterpret_cast<unsigned long long*>(order_all + 32 * WQ_CHUNKS * WQ_WARPS)[tid] = 0ULL;
    __syncthreads();
    const bool grp_lead = WQ_GROUP_STRIDED ? (wid < NGROUPS) : (wid % WQ_GROUP == 0);
#endif
#if WQ_SYNC || WQ_ALOAD || WQ_FULL_RECOUNT
    for (;; round++) {
        // class populations of this warp's pool: every lane adds a one-hot byte per owned slot (class c -> byte c of a
        // 64-bit word; retired slots land in the unused byte 6), two REDUX adds sum the halves over the warp
        uint32_t cls[WQ_K];
#if WQ_COUNT64
        unsigned long long oh = 0;
#pragma unroll
        for (int k = 0; k < WQ_K; k++) {
            cls[k] = S.state[base + 32 * k + lane] & 0xffu;
            oh += 1ULL << (8 * cls[k]);
        }
        const uint32_t pa = __reduce_add_sync(0xffffffffu, (uint32_t)oh);            // LOAD, STEP, COULOMB, RBEB
        const uint32_t pb = __reduce_add_sync(0xffffffffu, (uint32_t)(oh >> 32)) & 0xffffu;   // LOADWAIT, OTHER
#else
        uint32_t pa = 0, pb = 0;
#pragma unroll
        for (int k = 0; k < WQ_K; k++) {
            cls[k] = S.state[base + 32 * k + lane] & 0xffu;
            pa += cls[k] < 4u ? (1u << (8 * cls[k])) : 0u;                          // LOAD, STEP, COULOMB, RBEB
            pb += (cls[k] == WS_LOADWAIT) ? 1u : (cls[k] == WS_OTHER ? 256u : 0u);
        }
        pa = __reduce_add_sync(0xffffffffu, pa);
        pb = __reduce_add_sync(0xffffffffu, pb);
#endif
        // lane c: population of class c in this warp's pool, capped at one chunk
        const int n_c = lane < WS_IDLE ? (int)(((lane < 4 ? pa : pb) >> (8 * (lane & 3))) & 0xffu) : 0;
        const int f_c = n_c < 32 ? n_c : 32;
        int best;
#if WQ_SYNC
        {
            const unsigned cb3 = round % 3;
            // the six capped populations as 10-bit fields of one 64-bit word: low half = classes 0-2, high half = 3-5
            const unsigned fld = (unsigned)f_c << (10 * (myc % 3));
            const unsigned wlo = __reduce_add_sync(0xffffffffu, lane < 3 ? fld : 0u);
            const unsigned whi = __reduce_add_sync(0xffffffffu, (lane >= 3 && lane < WS_IDLE) ? fld : 0u);
            if (lane == 0 && (wlo | whi) != 0u) atomicAdd(csum + cb3, ((unsigned long long)whi << 32) | wlo);
            if (NGROUPS == 1) __syncthreads();
            else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(32 * WQ_GROUP) : "memory");
            const unsigned long long tot = csum[cb3];
            if (grp_lead && lane == 0) csum[(cb3 + 2) % 3] = 0ULL;      // read two barriers ago, added to after the next barrier
            if (tot == 0ULL) break;                                     // CTA-uniform: every slot of every warp is retired
            // the CTA's top class: most runnable entries; a class that waited WQ_AGE_FORCE rounds goes first
            const int N_c = lane < WS_IDLE ? (int)(((unsigned)(tot >> (32 * (myc / 3))) >> (10 * (myc % 3))) & 1023u) : 0;
            const int tsc = N_c > 0 ? N_c + (age >= WQ_AGE_FORCE ? 4096 + (age << 8) : 0) : 0;
            const unsigned tkey = __reduce_max_sync(0xffffffffu, ((unsigned)tsc << 3) | (unsigned)(7 - myc));
            const int top = 7 - (int)(tkey & 7u);
            const bool forced = (tkey >> 3) >= 4096u;
            age = (lane == top || N_c == 0) ? 0 : (age < 31 ? age + 1 : 31);
            // this warp: follow the CTA unless the own pool is clearly better served by another class
            const unsigned okey = __reduce_max_sync(0xffffffffu, ((unsigned)f_c << 3) | (unsigned)(7 - myc));
            const int oscore = (int)(okey >> 3);
            if (oscore == 0) continue;                                  // nothing pending here: keep the barrier count, wait
            const int ntop = __shfl_sync(0xffffffffu, f_c, top);
            best = (ntop > 0 && (forced || ntop + WQ_FOLLOW_SLACK >= oscore)) ? top : 7 - (int)(okey & 7u);
        }
#else
        {
            if ((pa | pb) == 0u) break;                                 // every slot of this warp is retired
            // most pending entries (capped at a full chunk) plus the age bonus; ties -> lower class
            // (LOAD / LOADWAIT units are short: they may run with fewer lanes, so that finished slots do not idle)
            const int wgt = (lane == WS_LOAD || lane == WS_LOADWAIT) ? WQ_LOAD_WEIGHT : 8;
            const int sc = n_c > 0 ? f_c * wgt + (age << WQ_AGE_SHIFT) : 0;
            const unsigned key = __reduce_max_sync(0xffffffffu, ((unsigned)sc << 3) | (unsigned)(7 - myc));
            best = 7 - (int)(key & 7u);
            age = (lane == best || n_c == 0) ? 0 : (age < 31 ? age + 1 : 31);
        }
#endif
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
        if (WQ_ALOAD && best == WS_LOADWAIT) {        // rows requested by ANY lane of this warp have landed and are visible to all
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
        }
        if (has) {
            const int it = (int)order[lane];
            const uint32_t sw = S.state[it];
            wf_execute_unit<SP, TK, FIRST, CB, WQ_ALOAD != 0>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, amask, lane, ltmask, row_counter, i0, i1, nsub, rows, -1,
                                                              WQ_OTHER_OUTLINE ? &wq_other_unit<SP, TK, FS> : nullptr,
                                                          (WQ_LOAD_OUTLINE && !WQ_ALOAD) ? &wq_load_unit<SP, TK, FIRST, FS> : nullptr);
        }
        __syncwarp();
    }

#else
    // Lean round (WQ_FULL_RECOUNT=0; NOT the default — measured slower).  The SASS page of the ncu capture of the round
    // above (profiles/r2_bench_main_kernel_ncu_summary.csv) counts 181 scheduler instructions per round next to 475 for a STEP
    // unit, 20.7 % of everything the kernel executes.  This version keeps the class populations INCREMENTALLY in two
    // warp-uniform words (after the unit every lane re-reads the state word of the slot it executed, two REDUX adds, the
    // executed class loses `cnt`), rotates the start group with a running offset and enters the unit through a warp-uniform
    // switch: ~165 instructions per round in SASS (the compiler spills the loop-carried words and rematerialises the lane
    // constants), and on B200 the main pass of 4e6 electrons took 30.3 ms against 27.5 ms (two runs each, same box,
    // gpurun_out/r2h1.log): the re-read of the executed slot + REDUX sits on the warp's critical path between two units,
    // and with four warps per scheduler the kernel is bound by per-warp latency, not by instruction count.  Taking further
    // chunks of the same class in one round (WQ_CHUNKS = 2 / 3) was slower still (28.5 / 28.6 ms against 27.5 ms at equal
    // RBEB trial count).  Kept buildable; results are bit-identical.
    uint32_t pa = (uint32_t)WQ_NS, pb = 0u;     // byte-packed populations: pa = LOAD, STEP, COULOMB, RBEB; pb = LOADWAIT, OTHER
    int rot = 0;
    for (;;) {
        if ((pa | pb) == 0u) break;                                     // every slot of this warp is retired
        // lane c owns class c: most pending entries (capped at a full chunk) plus the age bonus; ties -> lower class
        const int n_c = lane < WS_IDLE ? (int)(((lane < 4 ? pa : pb) >> (8 * (lane & 3))) & 0xffu) : 0;
        const int f_c = n_c < 32 * WQ_CHUNKS ? n_c : 32 * WQ_CHUNKS;
        const int wgt = (lane == WS_LOAD || lane == WS_LOADWAIT) ? WQ_LOAD_WEIGHT : 8;
        const int sc = n_c > 0 ? f_c * wgt + (age << WQ_AGE_SHIFT) : 0;
        const unsigned key = __reduce_max_sync(0xffffffffu, ((unsigned)sc << 3) | (unsigned)(7 - myc));
        const int best = 7 - (int)(key & 7u);
        age = (lane == best || n_c == 0) ? 0 : (age < 31 ? age + 1 : 31);
        const int nbest = (int)((((best < 4) ? pa : pb) >> (8 * (best & 3))) & 0xffu);
        // entries taken this round: one chunk of up to 32, or (WQ_CHUNKS > 1) further chunks of the same class as long as
        // each of them has at least WQ_CHUNK_MIN entries — the scheduling decision and the unit's code are reused
        int cnt = nbest < 32 ? nbest : 32;
#if WQ_CHUNKS > 1
#pragma unroll
        for (int c = 1; c < WQ_CHUNKS; c++) if (nbest >= 32 * c + WQ_CHUNK_MIN) cnt = nbest < 32 * (c + 1) ? nbest : 32 * (c + 1);
#endif
        // compact the slot ids of the picked class; the starting group rotates so that no slot waits forever when its
        // class stays above the number of entries a round takes
        int pos = 0, o = rot;
#pragma unroll
        for (int kk = 0; kk < WQ_K; kk++) {
            const int sl = base + o + lane;
            const bool mine = (S.state[sl] & 0xffu) == (uint32_t)best;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            const int q = pos + __popc(m & ltmask);
            if (mine && q < 32 * WQ_CHUNKS) order[q] = (unsigned short)sl;
            pos += __popc(m);
            o = (o + 32 == WQ_NS) ? 0 : o + 32;
        }
        rot = (rot + 32 == WQ_NS) ? 0 : rot + 32;
        __syncwarp();
        uint32_t oa = 0u, ob = 0u;                                      // one-hot sums of the classes the executed slots end up in
#if WQ_CHUNKS > 1
#pragma unroll 1
        for (int c0 = 0; c0 < cnt; c0 += 32)
#else
        const int c0 = 0;
#endif
        {
            const int left = cnt - c0;
            const bool has = lane < left;
            const unsigned amask = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
            if (has) {
                const int it = (int)order[c0 + lane];
                const uint32_t sw = S.state[it];
                wf_execute_unit<SP, TK, FIRST, CB, false>(P, T, Q, S, TS, tcum, fastsel, rc, cut, it, sw, amask, lane, ltmask, row_counter, i0, i1, nsub, rows, best,
                                                          WQ_OTHER_OUTLINE ? &wq_other_unit<SP, TK, FS> : nullptr,
                                                          (WQ_LOAD_OUTLINE && !WQ_ALOAD) ? &wq_load_unit<SP, TK, FIRST, FS> : nullptr);
                const uint32_t ns = S.state[it] & 0xffu;                // class of the slot after the unit
                oa += ns < 4u ? (1u << (8 * ns)) : 0u;
                ob += ns == (uint32_t)WS_LOADWAIT ? 1u : (ns == (uint32_t)WS_OTHER ? 256u : 0u);
            }
            __syncwarp();
        }
        pa += __reduce_add_sync(0xffffffffu, oa);
        pb += __reduce_add_sync(0xffffffffu, ob);
        if (best < 4) pa -= (uint32_t)cnt << (8 * best); else pb -= (uint32_t)cnt << (8 * (best - 4));
#ifdef WQ_TRACE
        round++;
#endif
    }
#endif

    for (int off = 16; off > 0; off >>= 1) nsub += __shfl_down_sync(0xffffffffu, nsub, off);
    if (lane == 0 && nsub) atomicAdd(P.substeps + SP, nsub);
    if (lane == 0 && round) { atomicMax(P.dbg, (unsigned long long)round); atomicAdd(P.dbg + 1, (unsigned long long)round); atomicAdd(P.dbg + 2, 1ULL); }
}

}  // namespace ptl
