// ptl_launch.cuh — host-side launch logic of the advance kernels (kernel choice, grid sizing, shared-memory opt-in).
// Included by ptl_adv_species.cu, which instantiates launch_advance_s for ONE species per translation unit.
#pragma once
#include "ptl_host.h"
#include "ptl_advance.cuh"
#include "ptl_advance_wf.cuh"
#include "ptl_advance_bq.cuh"
#include "ptl_advance_wq.cuh"

#ifndef PTL_DEFAULT_LEPTON_KERNEL
#define PTL_DEFAULT_LEPTON_KERNEL 5     // 3 = bq (list-scheduled, round 1), 5 = wq (warp-private pools, round 2)
#endif

#ifndef WQ_USE_FS
#define WQ_USE_FS 0                     // 1: tables of the usual shape run a kernel with the fast selection compiled in (template parameter FS);
                                        // measured 9 % SLOWER (27.6 against 25.3 ms): fewer instructions, but the layout ptxas then picks misses more in the instruction cache
#endif

namespace ptl_host {

template <int SP, bool FIRST, bool CB>
int32_t launch_advance_t(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, size_t smem, const long long* rows = nullptr) {
    auto kern = k_advance<SP, FIRST, CB>;
    // Function attributes are per DEVICE and this library may hold contexts on several devices in one process (and be
    // driven from several host threads): set the opt-in on every launch (a cheap driver call) instead of once per process.
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int blocks_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, ADV_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        blocks_per_sm = 1;
    // lanes per warp that carry a particle: as few as still give every resident warp of the machine something to do
    // (index-list mode does not know its row count on the host: full warps)
    const long long machine_warps = (long long)ctx->sm_count * blocks_per_sm * (ADV_THREADS / 32);
    int rpw = 32;
    if (rows == nullptr) {
        long long q = (i1 - i0 + machine_warps - 1) / machine_warps;
        rpw = q < 1 ? 1 : (q > 32 ? 32 : (int)q);
    }
    long long tiles = (i1 - i0 + rpw - 1) / rpw;
    long long want = (tiles + (ADV_THREADS / 32) - 1) / (ADV_THREADS / 32);
    long long grid = (long long)ctx->sm_count * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    CK(cudaMemsetAsync(&ctx->d_sc->tile_counter[SP], 0, sizeof(unsigned long long), ctx->lstream));
    bool timed = ctx->profiling && (i1 - i0) > ctx->stats.main_rows;
    if (timed) { ctx->stats.main_rows = i1 - i0; cudaEventRecord(ctx->ev0, ctx->lstream); }
    kern<<<(unsigned)grid, ADV_THREADS, smem, ctx->lstream>>>(A, i0, i1, &ctx->d_sc->tile_counter[SP], rows, rows ? &ctx->d_sc->slow_count[SP] : nullptr, rpw);
    if (timed) { cudaEventRecord(ctx->ev1, ctx->lstream); ctx->ev_pending = true; }
    LAUNCHED();
    ctx->stats.launches++;
    return 0;
}

// wavefront variant (collision-dominated species): persistent CTAs, shared-memory particle pool
template <int SP, int TK, bool FIRST, bool CB>
int32_t launch_advance_wf_k(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, size_t table_smem, const long long* rows) {
    auto kern = k_advance_wf<SP, TK, FIRST, CB>;
    const TableView& TV = A.tab[SP];
    size_t tsm = sizeof(ptl_process_desc) * TV.nprocs;
    if (TV.kind == 0) tsm += sizeof(double) * ((size_t)((TV.order == 3 && TV.nprocs <= 16) ? WF_CUM_STRIDE : TV.order * TV.nprocs) * (TV.k + 1) + (size_t)TV.order * (TV.k + 1));
    (void)table_smem;
    size_t smem = wf_pool_bytes() + tsm + 32;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);   // per device: see launch_advance_t
    int blocks_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, WF_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        blocks_per_sm = 1;
    long long nrow = i1 - i0;
    long long want = (nrow + WF_THREADS - 1) / WF_THREADS;
    long long grid = (long long)ctx->sm_count * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    CK(cudaMemsetAsync(&ctx->d_sc->tile_counter[SP], 0, sizeof(unsigned long long), ctx->lstream));
    bool timed = ctx->profiling && (i1 - i0) > ctx->stats.main_rows;
    if (timed) { ctx->stats.main_rows = i1 - i0; cudaEventRecord(ctx->ev0, ctx->lstream); }
    kern<<<(unsigned)grid, WF_THREADS, smem, ctx->lstream>>>(A, i0, i1, &ctx->d_sc->tile_counter[SP], rows, rows ? &ctx->d_sc->slow_count[SP] : nullptr);
    if (timed) { cudaEventRecord(ctx->ev1, ctx->lstream); ctx->ev_pending = true; }
    LAUNCHED();
    ctx->stats.launches++;
    return 0;
}

// list-scheduled variant (incremental per-class lists, one barrier per round, two chunks per warp)
template <int SP, int TK, bool FIRST, bool CB>
int32_t launch_advance_bq_k(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, const long long* rows) {
    auto kern = k_advance_bq<SP, TK, FIRST, CB>;
    const TableView& TV = A.tab[SP];
    size_t tsm = sizeof(ptl_process_desc) * TV.nprocs;
    if (TV.kind == 0) tsm += sizeof(double) * ((size_t)((TV.order == 3 && TV.nprocs <= 16) ? WF_CUM_STRIDE : TV.order * TV.nprocs) * (TV.k + 1) + (size_t)TV.order * (TV.k + 1));
    size_t smem = BQ_POOL_BYTES + tsm + 32;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);   // per device: see launch_advance_t
    int blocks_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, BQ_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        blocks_per_sm = 1;
    long long nrow = i1 - i0;
    long long want = (nrow + BQ_SLOTS - 1) / BQ_SLOTS;
    long long grid = (long long)ctx->sm_count * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    CK(cudaMemsetAsync(&ctx->d_sc->tile_counter[SP], 0, sizeof(unsigned long long), ctx->lstream));
    bool timed = ctx->profiling && (i1 - i0) > ctx->stats.main_rows;
    if (timed) { ctx->stats.main_rows = i1 - i0; cudaEventRecord(ctx->ev0, ctx->lstream); }
    kern<<<(unsigned)grid, BQ_THREADS, smem, ctx->lstream>>>(A, i0, i1, &ctx->d_sc->tile_counter[SP], rows, rows ? &ctx->d_sc->slow_count[SP] : nullptr);
    if (timed) { cudaEventRecord(ctx->ev1, ctx->lstream); ctx->ev_pending = true; }
    LAUNCHED();
    ctx->stats.launches++;
    return 0;
}

// warp-private variant (round 2): every warp owns WQ_NS slots; no CTA barrier, no shared lists
template <int SP, int TK, bool FIRST, bool CB, int FS = 0>
int32_t launch_advance_wq_k(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, const long long* rows) {
    auto kern = k_advance_wq<SP, TK, FIRST, CB, FS>;
    const TableView& TV = A.tab[SP];
    size_t tsm = sizeof(ptl_process_desc) * TV.nprocs;
    if (TV.kind == 0) tsm += sizeof(double) * ((size_t)((TV.order == 3 && TV.nprocs <= 16) ? WF_CUM_STRIDE : TV.order * TV.nprocs) * (TV.k + 1) + (size_t)TV.order * (TV.k + 1));
    size_t smem = WQ_POOL_BYTES + tsm + 32;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    // per device: see launch_advance_t
    int blocks_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, WQ_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        blocks_per_sm = 1;
    // Small passes (newborns, the reference's 1e4-electron swarms) are bound by the sequential chain of one particle,
    // not by throughput: spread the rows over as many warps as have at least one full chunk of them.
    long long nrow = i1 - i0;
    long long want = (nrow + 32 * WQ_WARPS - 1) / (32 * WQ_WARPS);
    long long grid = (long long)ctx->sm_count * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    CK(cudaMemsetAsync(&ctx->d_sc->tile_counter[SP], 0, sizeof(unsigned long long), ctx->lstream));
    bool timed = ctx->profiling && (i1 - i0) > ctx->stats.main_rows;
    if (timed) { ctx->stats.main_rows = i1 - i0; cudaEventRecord(ctx->ev0, ctx->lstream); }
    kern<<<(unsigned)grid, WQ_THREADS, smem, ctx->lstream>>>(A, i0, i1, &ctx->d_sc->tile_counter[SP], rows, rows ? &ctx->d_sc->slow_count[SP] : nullptr);
    if (timed) { cudaEventRecord(ctx->ev1, ctx->lstream); ctx->ev_pending = true; }
    LAUNCHED();
    ctx->stats.launches++;
    return 0;
}

template <int SP, bool FIRST, bool CB>
int32_t launch_advance_wf_t(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, size_t table_smem, const long long* rows = nullptr) {
    // PTL_KERNEL=wq: warp-private pools (k_advance_wq); bq: list-scheduled kernel (k_advance_bq); wf: the re-sorting kernel
    // — kept for A/B measurements and run against each other by the parity tests.
    const int variant = ctx->lepton_kernel ? ctx->lepton_kernel : PTL_DEFAULT_LEPTON_KERNEL;
    if (variant == 5) {
        if (A.tab[SP].kind == 0) {
            // the usual table shape (Chebyshev, order 3, <= 16 processes) gets a kernel with the fast selection compiled in
            if (WQ_USE_FS && A.tab[SP].order == 3 && A.tab[SP].nprocs <= 16) return launch_advance_wq_k<SP, 0, FIRST, CB, WQ_USE_FS>(ctx, A, i0, i1, rows);
            return launch_advance_wq_k<SP, 0, FIRST, CB, 0>(ctx, A, i0, i1, rows);
        }
        return launch_advance_wq_k<SP, 1, FIRST, CB>(ctx, A, i0, i1, rows);
    }
    if (variant != 4) {
        if (A.tab[SP].kind == 0) return launch_advance_bq_k<SP, 0, FIRST, CB>(ctx, A, i0, i1, rows);
        return launch_advance_bq_k<SP, 1, FIRST, CB>(ctx, A, i0, i1, rows);
    }
    if (A.tab[SP].kind == 0) return launch_advance_wf_k<SP, 0, FIRST, CB>(ctx, A, i0, i1, table_smem, rows);
    return launch_advance_wf_k<SP, 1, FIRST, CB>(ctx, A, i0, i1, table_smem, rows);
}

// First-pass kernel choice.  Photons, and any species whose measured kappa (sub-steps per row of the previous advance)
// is small, are HBM-bound: free flights go through the streaming kernel and only the rows that collide within dt are
// deferred, through an index list, to the general kernel of the species (one particle per lane for photons, wavefront
// for leptons).  Collision-dominated populations go straight to the wavefront kernel.
template <int SP>
int32_t launch_advance_s(ptl_context* ctx, const AdvanceParams& A, long long i0, long long i1, bool first, bool cb, size_t smem, bool low_kappa) {
    const long long* rows = nullptr;
    if ((SP == PTL_PHOTON || low_kappa) && !cb && (i0 & 1) == 0 && i1 - i0 >= 4096 && ctx->use_stream) {
        size_t need = (size_t)(i1 - i0);
        if (need > ctx->slow_cap[SP]) {
            cudaFree(ctx->d_slow_rows[SP]);
            ctx->d_slow_rows[SP] = nullptr; ctx->slow_cap[SP] = 0;
            // grow geometrically (x2, at least 4 Mi entries): a photon population that grows every step must not pay a
            // cudaFree/cudaMalloc pair inside most advance! calls (each one synchronises the device)
            size_t cap = need * 2 > ((size_t)4 << 20) ? need * 2 : ((size_t)4 << 20);
            CK(cudaMalloc(&ctx->d_slow_rows[SP], sizeof(long long) * cap));
            ctx->slow_cap[SP] = cap;
        }
        CK(cudaMemsetAsync(&ctx->d_sc->slow_count[SP], 0, sizeof(unsigned long long), ctx->lstream));
        const TableView& TV = A.tab[SP];
        size_t ssm = TV.kind == 0 ? sizeof(double) * TV.order * (TV.k + 1) : 8;
        constexpr int np = (SP == PTL_PHOTON) ? 2 : 1;      // particles per thread (k_advance_stream)
        long long pairs = (i1 - i0 + np - 1) / np;
        long long grid = (pairs + STREAM_THREADS - 1) / STREAM_THREADS;
        long long maxgrid = (long long)ctx->sm_count * 8;
        if (grid > maxgrid) grid = maxgrid;
        bool timed = ctx->profiling && (i1 - i0) > ctx->stats.main_rows;
        if (timed) { ctx->stats.main_rows = i1 - i0; cudaEventRecord(ctx->ev0, ctx->lstream); }
        bool tma = false;
        if constexpr (SP != PTL_PHOTON) {
            // leptons: rows staged through shared memory by bulk copies (k_advance_stream_tma); needs 16-byte aligned tiles
            // of the byte column too, i.e. a pass that starts on a multiple of 16 rows (the first pass always does)
            tma = ctx->use_stream_tma && (i0 & 15) == 0;
            if (tma) {
                const size_t tsm = STT_RING_BYTES + ssm;
                auto k1 = k_advance_stream_tma<SP, true>;
                auto k0 = k_advance_stream_tma<SP, false>;
                cudaFuncSetAttribute(first ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);   // per device, see launch_advance_t
                long long tiles = (i1 - i0 + STT_ROWS - 1) / STT_ROWS;
                long long tg = (long long)ctx->sm_count * STT_MINB;
                if (tg > tiles) tg = tiles;
                if (first) k1<<<(unsigned)tg, STT_THREADS, tsm, ctx->lstream>>>(A, i0, i1, ctx->d_slow_rows[SP], &ctx->d_sc->slow_count[SP]);
                else k0<<<(unsigned)tg, STT_THREADS, tsm, ctx->lstream>>>(A, i0, i1, ctx->d_slow_rows[SP], &ctx->d_sc->slow_count[SP]);
            }
        }
        if (tma) {
        } else if (first) k_advance_stream<SP, true><<<(unsigned)grid, STREAM_THREADS, ssm, ctx->lstream>>>(A, i0, i1, ctx->d_slow_rows[SP], &ctx->d_sc->slow_count[SP]);
        else k_advance_stream<SP, false><<<(unsigned)grid, STREAM_THREADS, ssm, ctx->lstream>>>(A, i0, i1, ctx->d_slow_rows[SP], &ctx->d_sc->slow_count[SP]);
        if (timed) { cudaEventRecord(ctx->ev1, ctx->lstream); ctx->ev_pending = true; }
        LAUNCHED();
        ctx->stats.launches++;
        rows = ctx->d_slow_rows[SP];
        cb = false;
    }
    // Small passes (the newborns of a step, the reference's own 1e4-electron swarms) are bound by the sequential chain of
    // ONE particle (up to ~400 collisions within dt), not by throughput: the one-particle-per-lane kernel keeps the state
    // in registers and has no scheduling rounds, so a chain runs ~2.5x faster there; the wavefront kernels win as soon as
    // there are enough rows to fill the machine (ctx->small_pass_rows, ptl_set_option "small_pass_rows").
    const bool small_pass = SP != PTL_PHOTON && rows == nullptr && (i1 - i0) < ctx->small_pass_rows;
    if constexpr (SP != PTL_PHOTON) {
        if (!small_pass) {
            if (first) return cb ? launch_advance_wf_t<SP, true, true>(ctx, A, i0, i1, smem, rows) : launch_advance_wf_t<SP, true, false>(ctx, A, i0, i1, smem, rows);
            return cb ? launch_advance_wf_t<SP, false, true>(ctx, A, i0, i1, smem, rows) : launch_advance_wf_t<SP, false, false>(ctx, A, i0, i1, smem, rows);
        }
    }
    if (first) return cb ? launch_advance_t<SP, true, true>(ctx, A, i0, i1, smem, rows) : launch_advance_t<SP, true, false>(ctx, A, i0, i1, smem, rows);
    return cb ? launch_advance_t<SP, false, true>(ctx, A, i0, i1, smem, rows) : launch_advance_t<SP, false, false>(ctx, A, i0, i1, smem, rows);
}

}  // namespace ptl_host
