// ptl_host.h — host-side state behind the C ABI (opaque to callers): the context, its tables / populations / scratch, and the
// small helpers every translation unit of the library shares.  The library is split into one translation unit for the ABI
// (ptl_api.cu) and one per species for the advance kernels (ptl_adv_species.cu, compiled four times) so that the
// sm_100a build runs in parallel.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include "ptl_common.cuh"

namespace ptl { constexpr int DIAG_NVAL_HOST = 12; constexpr int PTL_COLL_SCRATCH = 4096; }

using namespace ptl;

namespace ptl_host {

struct DeviceScalars {            // one small device block mirrored in pinned host memory
    int flags;
    int _pad;
    unsigned long long substeps[PTL_NSPECIES], births, tile_counter[PTL_NSPECIES], total, nmoves, slow_count[PTL_NSPECIES], max_uid;
    unsigned long long pop_n[64];
    unsigned long long wall_n[PTL_MAX_WALLS];
    double diag[ptl::DIAG_NVAL_HOST];
    unsigned long long dbg[64];    // PTL_TRACE: max / sum of scheduler rounds per CTA, CTA count
};

struct Table {
    TableView v{};
    std::vector<ptl_process_desc> procs;
    double *d_rate = nullptr, *d_rb = nullptr, *d_cum = nullptr, *d_cum2 = nullptr, *d_rbvec = nullptr;
    ptl_process_desc* d_procs = nullptr;
    unsigned long long* d_counts = nullptr;
    size_t smem_bytes = 0;
};

struct Pop {
    PopView v{};
    void* block = nullptr;
    long long iup = 0;
    int table = -1;
    int slot = -1;                // index into DeviceScalars.pop_n
    double kappa_est = -1;        // measured sub-steps per row in the last advance (< 0: unknown)
    long long rows_last = 0;
    bool alive = false;
};

struct MultiPop {
    std::vector<int> pops;
    int by_species[PTL_NSPECIES];
};

struct Sb { SbView v{}; };
struct ChebLoss { ChebLossView v{}; };

struct Wall {
    WallBuf b{};
    void* block = nullptr;
};

}  // namespace ptl_host
using namespace ptl_host;

struct ptl_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<Table> tables;
    std::vector<Pop> pops;
    std::vector<Sb> sbs;
    std::vector<ChebLoss> cls;
    std::vector<MultiPop> mps;
    Wall walls[PTL_MAX_WALLS];
    DeviceScalars* d_sc = nullptr;
    DeviceScalars* h_sc = nullptr;     // pinned
    uint64_t seed = 0;
    uint32_t step = 0;
    uint64_t next_uid = 1;
    ptl_advance_stats stats{};
    std::string err;
    // scratch
    void* stage[2] = {nullptr, nullptr};
    size_t stage_rows = 0;
    unsigned int* d_tile_counts = nullptr;
    unsigned long long* d_tile_offsets = nullptr;
    size_t tiles_cap = 0;
    long long *d_holes = nullptr, *d_tails = nullptr;
    size_t moves_cap = 0;
    double* d_partial = nullptr;
    int partial_blocks = 0;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    long long* d_slow_rows[PTL_NSPECIES] = {nullptr, nullptr, nullptr, nullptr};  // rows the streaming kernel of a species deferred to its general kernel
    size_t slow_cap[PTL_NSPECIES] = {0, 0, 0, 0};
    // species of one pass of advance1! are independent (births only append beyond the rows a pass visits): their kernels run
    // concurrently, the first on the context's stream and the others on auxiliary streams forked from / joined to it
    cudaStream_t lstream = nullptr;    // stream the advance launchers use right now
    cudaStream_t aux[PTL_NSPECIES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[PTL_NSPECIES] = {nullptr, nullptr, nullptr, nullptr};
    bool overlap_species = true;       // ptl_set_option "overlap" (0: every kernel of a pass on one stream, in tuple order)
    int lepton_kernel = 0;             // 0 = default (PTL_DEFAULT_LEPTON_KERNEL), 3 = bq, 4 = wf, 5 = wq (ptl_set_option "kernel" / PTL_KERNEL)
    long long small_pass_rows = 16384; // lepton passes with fewer rows run on the one-particle-per-lane kernel (chain latency, not throughput)
    bool use_stream = true;            // streaming fast path for low-kappa species (ptl_set_option "stream" / PTL_KERNEL=nostream)
    bool use_stream_tma = false;       // leptons on the streaming path: TMA-staged kernel (ptl_set_option "stream_tma" / PTL_KERNEL=tma); measured slower, off
    long long launch_total = 0;        // kernels launched since the last ptl_launch_count(reset)
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_pending = false;
    cudaEvent_t ev_copy = nullptr;     // "the host arrays of this upload have been read"
    // multi-GPU (ptl_comm.cu): NCCL communicator bound at run time, this context's rank, a small device scratch
    void* comm = nullptr;
    int rank = 0, nranks = 1;
    double* d_coll = nullptr;
};

namespace ptl_host {

inline bool cuda_ok(ptl_context* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define CK(call) do { if (!cuda_ok(ctx, (call), #call)) return PTL_ECUDA; } while (0)
#define LAUNCHED() do { ctx->launch_total++; CK(cudaGetLastError()); } while (0)
// Every entry point binds the calling host thread to the context's device: a host thread that was not the one that created
// the context (the e2e leg of bench.py drives three contexts from three threads) starts on device 0, and on any other rank
// of a multi-GPU job every launch then failed with PTL_ECUDA.
#define PTL_BIND(c) do { if (c) cudaSetDevice((c)->device); } while (0)

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

inline int32_t sync_scalars(ptl_context* ctx) {
    CK(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

inline Pop* get_pop(ptl_context* ctx, int32_t pop) {
    if (!ctx || pop < 0 || pop >= (int)ctx->pops.size() || !ctx->pops[pop].alive) return nullptr;
    return &ctx->pops[pop];
}

inline unsigned long long* dev_n(ptl_context* ctx, const Pop& P) { return &ctx->d_sc->pop_n[P.slot]; }

// read popl.n from the device (clamped to capacity; overflow raises the sticky flag)
inline int32_t read_n(ptl_context* ctx, Pop& P, long long* out) {
    int32_t rc = sync_scalars(ctx);
    if (rc) return rc;
    long long n = (long long)ctx->h_sc->pop_n[P.slot];
    if (n > P.v.capacity) {
        n = P.v.capacity;
        unsigned long long nn = (unsigned long long)n;
        CK(cudaMemcpyAsync(dev_n(ctx, P), &nn, sizeof(nn), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *out = n;
    return 0;
}

inline int32_t set_n(ptl_context* ctx, Pop& P, long long n) {
    unsigned long long nn = (unsigned long long)n;
    ctx->h_sc->pop_n[P.slot] = nn;
    CK(cudaMemcpyAsync(dev_n(ctx, P), &ctx->h_sc->pop_n[P.slot], sizeof(nn), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// Two staging buffers for the xyz-interleaved host vectors.  They hold `want` rows each when that is at most 16 Mi rows (384 MiB
// per buffer): an upload of that size then needs no buffer reuse, so its host->device copies do not wait for any kernel and the
// call can hand the borrowed host arrays back as soon as the COPIES are done (see ptl_population_upload).  Larger transfers go
// through 4 Mi-row chunks that alternate between the buffers.
inline int32_t ensure_stage(ptl_context* ctx, size_t want = 0) {
    const size_t chunk = (size_t)1 << 22, big = (size_t)1 << 24;
    size_t rows = want <= chunk ? chunk : (want <= big ? want : chunk);
    if (ctx->stage_rows >= rows) return 0;
    for (int b = 0; b < 2; b++) {
        if (ctx->stage[b]) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->stage[b]); ctx->stage[b] = nullptr; }
        CK(cudaMalloc(&ctx->stage[b], rows * 3 * sizeof(double)));
    }
    ctx->stage_rows = rows;
    return 0;
}

inline int32_t ensure_tmp(ptl_context* ctx, size_t bytes) {
    if (ctx->tmp_bytes >= bytes) return 0;
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    ctx->d_tmp = nullptr; ctx->tmp_bytes = 0;
    CK(cudaMalloc(&ctx->d_tmp, bytes));
    ctx->tmp_bytes = bytes;
    return 0;
}

inline void fill_params(ptl_context* ctx, const MultiPop* mp, AdvanceParams& A) {
    memset(&A, 0, sizeof(A));
    for (int s = 0; s < PTL_NSPECIES; s++) {
        int pi = mp ? mp->by_species[s] : -1;
        if (pi >= 0) {
            A.pop[s] = ctx->pops[pi].v;
            A.pop[s].present = 1;
            A.tab[s] = ctx->tables[ctx->pops[pi].table].v;
        }
    }
    for (size_t i = 0; i < ctx->sbs.size() && i < (size_t)MAX_SB; i++) A.sb[i] = ctx->sbs[i].v;
    for (size_t i = 0; i < ctx->cls.size() && i < (size_t)MAX_CHEBLOSS; i++) A.cl[i] = ctx->cls[i].v;
    for (int k = 0; k < PTL_MAX_WALLS; k++) A.wall[k] = ctx->walls[k].b;
    A.seed_lo = (uint32_t)ctx->seed;
    A.seed_hi = (uint32_t)(ctx->seed >> 32);
    A.step = ctx->step;
    A.flags = &ctx->d_sc->flags;
    A.substeps = ctx->d_sc->substeps;
    A.births = &ctx->d_sc->births;
    A.dbg = ctx->d_sc->dbg;
}

// local halves of the diagnostics, shared by ptl_diag / ptl_histogram (ptl_api.cu) and their all-reduce forms (ptl_comm.cu)
int32_t diag_local_launch(ptl_context* ctx, Pop& P, long long n);     // leaves the 12-vector in d_sc->diag ([11] = n)
void diag_unpack(const double* d, ptl_diag_out* out);
int32_t histogram_local_launch(ptl_context* ctx, Pop& P, int32_t quantity, double lo, double hi, int32_t nbins, int32_t logscale);   // bins in d_tmp

// advance launchers: one explicit instantiation per species, each in its own translation unit (ptl_adv_species.cu)
template <int SP>
int32_t launch_advance_s(ptl_context* ctx, const ptl::AdvanceParams& A, long long i0, long long i1, bool first, bool cb, size_t smem, bool low_kappa);
extern template int32_t launch_advance_s<PTL_ELECTRON>(ptl_context*, const ptl::AdvanceParams&, long long, long long, bool, bool, size_t, bool);
extern template int32_t launch_advance_s<PTL_PHOTON>(ptl_context*, const ptl::AdvanceParams&, long long, long long, bool, bool, size_t, bool);
extern template int32_t launch_advance_s<PTL_POSITRON>(ptl_context*, const ptl::AdvanceParams&, long long, long long, bool, bool, size_t, bool);
extern template int32_t launch_advance_s<PTL_SLOW_ELECTRON>(ptl_context*, const ptl::AdvanceParams&, long long, long long, bool, bool, size_t, bool);

}  // namespace ptl_host
