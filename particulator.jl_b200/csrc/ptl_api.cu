// ptl_api.cu — the C ABI of include/particulator_b200.h: context, device-resident stores, table
// upload, and the host-side launch logic of every kernel (advance pass loop, compaction, diagnostics).
// Plain C signatures only; no torch / C++ types cross the boundary.  There is NO CPU fallback: a
// context can only be created on a compute-capability-10.x device.
#include <cub/device/device_radix_sort.cuh>

#include "ptl_host.h"
#include "ptl_advance.cuh"
#include "ptl_store.cuh"

using namespace ptl;
using namespace ptl_host;
static_assert(ptl::DIAG_NVAL == ptl::DIAG_NVAL_HOST, "diagnostic vector length");

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {

int32_t launch_advance(ptl_context* ctx, int species, const AdvanceParams& A, long long i0, long long i1, bool first, bool cb, size_t smem, bool low_kappa) {
    switch (species) {
    case PTL_ELECTRON: return launch_advance_s<PTL_ELECTRON>(ctx, A, i0, i1, first, cb, smem, low_kappa);
    case PTL_PHOTON: return launch_advance_s<PTL_PHOTON>(ctx, A, i0, i1, first, cb, smem, low_kappa);
    case PTL_POSITRON: return launch_advance_s<PTL_POSITRON>(ctx, A, i0, i1, first, cb, smem, low_kappa);
    case PTL_SLOW_ELECTRON: return launch_advance_s<PTL_SLOW_ELECTRON>(ctx, A, i0, i1, first, cb, smem, low_kappa);
    }
    return PTL_EINVAL;
}

#define DISPATCH_SPECIES(sp, ...)                                                   \
    switch (sp) {                                                                   \
    case PTL_ELECTRON: { constexpr int SP = PTL_ELECTRON; __VA_ARGS__; break; }     \
    case PTL_PHOTON: { constexpr int SP = PTL_PHOTON; __VA_ARGS__; break; }         \
    case PTL_POSITRON: { constexpr int SP = PTL_POSITRON; __VA_ARGS__; break; }     \
    default: { constexpr int SP = PTL_SLOW_ELECTRON; __VA_ARGS__; break; }          \
    }

unsigned grid_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// fold the device-side maximum of the explicitly uploaded sequential uids into the host counter
int32_t refresh_next_uid(ptl_context* ctx) {
    int32_t rc = sync_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->max_uid >= ctx->next_uid) ctx->next_uid = ctx->h_sc->max_uid + 1;
    return 0;
}

}  // namespace

// =====================================================================================================
// context
// =====================================================================================================
EXPORT int32_t ptl_abi_version(void) { return PTL_ABI_VERSION; }

EXPORT int32_t ptl_context_create(int32_t device, void* stream, ptl_context** out) {
    if (!out) return PTL_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return PTL_ENODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PTL_ENODEVICE;
    if (prop.major != 10) return PTL_ENODEVICE;   // sm_100a cubin only: no fallback of any kind
    if (cudaSetDevice(device) != cudaSuccess) return PTL_ECUDA;
    ptl_context* ctx = new ptl_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* km = getenv("PTL_KERNEL")) {      // A/B measurements; ptl_set_option does the same per context
        ctx->lepton_kernel = !strcmp(km, "wq") ? 5 : (!strcmp(km, "bq") ? 3 : (!strcmp(km, "wf") ? 4 : 0));
        if (!strcmp(km, "nostream")) ctx->use_stream = false;
        if (!strcmp(km, "tma")) ctx->use_stream_tma = true;
    }
    if (const char* sp = getenv("PTL_SMALL_PASS")) ctx->small_pass_rows = atoll(sp);
    if (const char* ov = getenv("PTL_OVERLAP")) ctx->overlap_species = atoi(ov) != 0;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PTL_ECUDA; }
        ctx->own_stream = true;
    }
    if (cudaMalloc(&ctx->d_sc, sizeof(DeviceScalars)) != cudaSuccess || cudaMallocHost(&ctx->h_sc, sizeof(DeviceScalars)) != cudaSuccess) {
        delete ctx;
        return PTL_ENOMEM;
    }
    cudaMemsetAsync(ctx->d_sc, 0, sizeof(DeviceScalars), ctx->stream);
    memset(ctx->h_sc, 0, sizeof(DeviceScalars));
    ctx->lstream = ctx->stream;
    ctx->partial_blocks = ctx->sm_count * 8;
    if (cudaMalloc(&ctx->d_partial, sizeof(double) * DIAG_NVAL * ctx->partial_blocks) != cudaSuccess) { delete ctx; return PTL_ENOMEM; }
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return 0;
}

EXPORT int32_t ptl_context_destroy(ptl_context* ctx) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->tables) { cudaFree(t.d_rate); cudaFree(t.d_rb); cudaFree(t.d_cum); cudaFree(t.d_cum2); cudaFree(t.d_rbvec); cudaFree(t.d_procs); cudaFree(t.d_counts); }
    for (auto& p : ctx->pops) if (p.block) cudaFree(p.block);
    for (auto& s : ctx->sbs) { cudaFree((void*)s.v.log_energy); cudaFree((void*)s.v.data); }
    for (auto& c : ctx->cls) { cudaFree((void*)c.v.ec); cudaFree((void*)c.v.pc); }
    for (auto& w : ctx->walls) if (w.block) cudaFree(w.block);
    for (int b = 0; b < 2; b++) if (ctx->stage[b]) cudaFree(ctx->stage[b]);
    cudaFree(ctx->d_tile_counts); cudaFree(ctx->d_tile_offsets); cudaFree(ctx->d_holes); cudaFree(ctx->d_tails);
    cudaFree(ctx->d_partial); cudaFree(ctx->d_tmp); cudaFree(ctx->d_coll);
    for (int k = 0; k < PTL_NSPECIES; k++) {
        cudaFree(ctx->d_slow_rows[k]);
        if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
        if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    cudaFree(ctx->d_sc);
    cudaFreeHost(ctx->h_sc);
    if (ctx->ev0) { cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); }
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

EXPORT const char* ptl_last_error(ptl_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

EXPORT int32_t ptl_error_flags(ptl_context* ctx, int32_t clear) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    int32_t rc = sync_scalars(ctx);
    if (rc) return rc;
    int f = ctx->h_sc->flags;
    if (clear) { CK(cudaMemsetAsync(&ctx->d_sc->flags, 0, sizeof(int), ctx->stream)); }
    return f;
}

EXPORT int32_t ptl_synchronize(ptl_context* ctx) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

EXPORT int32_t ptl_set_rng(ptl_context* ctx, uint64_t seed, uint32_t step) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    ctx->seed = seed; ctx->step = step;
    return 0;
}
// Tuning / A-B knobs that are not part of the reference's surface.  "kernel": lepton advance kernel variant
// (0 = default, 3 = bq list-scheduled, 4 = wf re-sorting, 5 = wq warp-private); "stream": 0 disables the streaming fast path.
EXPORT int32_t ptl_set_option(ptl_context* ctx, const char* name, int64_t value) {
    PTL_BIND(ctx);
    if (!ctx || !name) return PTL_EINVAL;
    if (!strcmp(name, "kernel")) {
        if (value != 0 && value != 3 && value != 4 && value != 5) return PTL_EINVAL;
        ctx->lepton_kernel = (int)value;
        return 0;
    }
    if (!strcmp(name, "stream")) { ctx->use_stream = value != 0; return 0; }
    if (!strcmp(name, "stream_tma")) { ctx->use_stream_tma = value != 0; return 0; }
    if (!strcmp(name, "overlap")) { ctx->overlap_species = value != 0; return 0; }
    if (!strcmp(name, "small_pass_rows")) { if (value < 0) return PTL_EINVAL; ctx->small_pass_rows = value; return 0; }
    ctx->err = std::string("unknown option ") + name;
    return PTL_EINVAL;
}

// uid counter behind default uids (ptl_population_upload / ptl_population_append without explicit uids).  Part of the
// restart state: a restored run must not reissue a uid that is still alive, because uids key the RNG streams.
EXPORT int32_t ptl_set_uid_counter(ptl_context* ctx, uint64_t next_uid) {
    if (!ctx || next_uid == 0) return PTL_EINVAL;
    ctx->next_uid = next_uid;
    return 0;
}
EXPORT uint64_t ptl_get_uid_counter(ptl_context* ctx) {
    PTL_BIND(ctx);
    if (!ctx || refresh_next_uid(ctx)) return 0;
    return ctx->next_uid;
}

EXPORT int32_t ptl_get_rng(ptl_context* ctx, uint64_t* seed, uint32_t* step) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    if (seed) *seed = ctx->seed;
    if (step) *step = ctx->step;
    return 0;
}

// =====================================================================================================
// tables
// =====================================================================================================
namespace {
int32_t upload_doubles(ptl_context* ctx, const double* src, size_t n, double** dst) {
    CK(cudaMalloc(dst, sizeof(double) * (n ? n : 1)));
    if (n) CK(cudaMemcpyAsync(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int32_t upload_procs(ptl_context* ctx, Table& T, const ptl_process_desc* procs, int nprocs) {
    T.procs.assign(procs, procs + nprocs);
    for (auto& p : T.procs) {
        if (p.kind == PTL_PROC_COULOMB) p.par[1] = pow(p.par[0], -1.0 / 3.0);   // Z^(-1//3), relativistic_coulomb.jl:14
        if (p.kind == PTL_PROC_SELTZER && (p.aux < 0 || p.aux >= (int)ctx->sbs.size())) return PTL_EHANDLE;
    }
    CK(cudaMalloc(&T.d_procs, sizeof(ptl_process_desc) * (nprocs ? nprocs : 1)));
    if (nprocs) CK(cudaMemcpyAsync(T.d_procs, T.procs.data(), sizeof(ptl_process_desc) * nprocs, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMalloc(&T.d_counts, sizeof(unsigned long long) * (nprocs + 1)));
    CK(cudaMemsetAsync(T.d_counts, 0, sizeof(unsigned long long) * (nprocs + 1), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    T.v.procs = T.d_procs;
    T.v.counts = T.d_counts;
    T.v.nprocs = nprocs;
    return 0;
}
}  // namespace

EXPORT int32_t ptl_sb_table_create(ptl_context* ctx, int32_t ncum, int32_t nE, const double* log_energy, const double* data) {
    PTL_BIND(ctx);
    if (!ctx || ncum < 2 || nE < 2 || !log_energy || !data) return PTL_EINVAL;
    if (ctx->sbs.size() >= (size_t)MAX_SB) return PTL_ENOMEM;
    Sb s;
    s.v.ncum = ncum; s.v.nE = nE;
    double *dl = nullptr, *dd = nullptr;
    int32_t rc = upload_doubles(ctx, log_energy, nE, &dl); if (rc) return rc;
    rc = upload_doubles(ctx, data, (size_t)ncum * nE, &dd); if (rc) return rc;
    s.v.log_energy = dl; s.v.data = dd;
    ctx->sbs.push_back(s);
    return (int32_t)ctx->sbs.size() - 1;
}

EXPORT int32_t ptl_table_create_cheb(ptl_context* ctx, int32_t order, int32_t nprocs, int32_t k, double xmax, const double* rate,
                                     const double* ratebound, const ptl_process_desc* procs) {
    PTL_BIND(ctx);
    if (!ctx || order < 1 || order > MAX_ORDER || nprocs < 0 || nprocs > PTL_MAX_PROCS || k < 1 || !(xmax > 0)) return PTL_EINVAL;
    Table T;
    T.v.kind = 0; T.v.order = order; T.v.k = k; T.v.xmax = xmax; T.v.rxmax = 1.0 / xmax;
    int32_t rc = upload_doubles(ctx, rate, (size_t)order * nprocs * (k + 1), &T.d_rate); if (rc) return rc;
    rc = upload_doubles(ctx, ratebound, (size_t)order * (k + 1), &T.d_rb); if (rc) return rc;
    T.v.rate = T.d_rate; T.v.ratebound = T.d_rb;
    {   // running sums over the (sorted) processes: cum[m, j, i] = sum_{j' <= j} rate[m, j', i]
        std::vector<double> cum((size_t)order * nprocs * (k + 1));
        for (int i = 0; i <= k; i++)
            for (int m = 0; m < order; m++) {
                double acc = 0;
                for (int j = 0; j < nprocs; j++) {
                    size_t q = (size_t)m + (size_t)order * ((size_t)j + (size_t)nprocs * i);
                    acc += rate[q];
                    cum[q] = acc;
                }
            }
        rc = upload_doubles(ctx, cum.data(), cum.size(), &T.d_cum); if (rc) return rc;
        T.v.cum = T.d_cum;
    }
    // Intervals on which no fitted rate dips below zero (fits do near process thresholds): exact minimum of
    // a0 + a1 x + a2 (2 x^2 - 1) over [-1, 1].  There the running sums are non-decreasing in the process index and the
    // kernels may select by binary search; elsewhere they scan sequentially like the reference.
    T.v.mono_mask = 0;
    if (order == 3 && k + 1 <= 64) {
        for (int i = 0; i <= k; i++) {
            bool ok = true;
            for (int j = 0; j < nprocs && ok; j++) {
                const double* a = rate + (size_t)order * ((size_t)j + (size_t)nprocs * i);
                double mn = fmin(a[0] - a[1] + a[2], a[0] + a[1] + a[2]);
                if (a[2] > 0 && fabs(a[1]) <= 4 * a[2]) { double x = -a[1] / (4 * a[2]); mn = fmin(mn, a[0] + a[1] * x + a[2] * (2 * x * x - 1)); }
                ok = mn >= 0;      // NaN -> false
            }
            if (ok) T.v.mono_mask |= 1ULL << i;
        }
    }
    rc = upload_procs(ctx, T, procs, nprocs); if (rc) return rc;
    T.smem_bytes = sizeof(double) * ((size_t)order * nprocs * (k + 1) + (size_t)order * (k + 1)) + sizeof(ptl_process_desc) * nprocs;
    if (T.smem_bytes > 60 * 1024) { ctx->err = "Chebyshev table too large for shared memory"; return PTL_EINVAL; }
    ctx->tables.push_back(T);
    return (int32_t)ctx->tables.size() - 1;
}

EXPORT int32_t ptl_table_create_linear(ptl_context* ctx, int32_t grid_kind, double L1, double L2, int32_t nE, int32_t nprocs,
                                       const double* rate, double maxrate, const ptl_process_desc* procs) {
    PTL_BIND(ctx);
    if (!ctx || nE < 2 || nprocs < 0 || nprocs > PTL_MAX_PROCS || (grid_kind != 0 && grid_kind != 1)) return PTL_EINVAL;
    Table T;
    T.v.kind = 1; T.v.grid_kind = grid_kind; T.v.L1 = L1; T.v.L2 = L2; T.v.nE = nE; T.v.maxrate = maxrate;
    int32_t rc = upload_doubles(ctx, rate, (size_t)nprocs * nE, &T.d_rate); if (rc) return rc;
    T.v.rate = T.d_rate;
    {   // cum[j, e] = sum_{j' <= j} rate[j', e]
        std::vector<double> cum((size_t)nprocs * nE);
        for (int e = 0; e < nE; e++) {
            double acc = 0;
            for (int j = 0; j < nprocs; j++) { acc += rate[j + (size_t)nprocs * e]; cum[j + (size_t)nprocs * e] = acc; }
        }
        rc = upload_doubles(ctx, cum.data(), cum.size(), &T.d_cum); if (rc) return rc;
        T.v.cum = T.d_cum;
        // the two grid rows a lookup interpolates between, side by side: a search probe is ONE 16-byte load (one sector)
        // instead of two 8-byte loads from rows np * 8 bytes apart (the LXCat kernel is bound by that sector traffic)
        std::vector<double> pair((size_t)2 * nprocs * nE);
        for (int e = 0; e < nE; e++)
            for (int j = 0; j < nprocs; j++) {
                size_t q = (size_t)j + (size_t)nprocs * e;
                pair[2 * q] = cum[q];
                pair[2 * q + 1] = (e + 1 < nE) ? cum[q + nprocs] : cum[q];
            }
        rc = upload_doubles(ctx, pair.data(), pair.size(), &T.d_cum2); if (rc) return rc;
        T.v.cum2 = reinterpret_cast<const double2*>(T.d_cum2);
    }
    {   // binary-search selection needs non-decreasing running sums: every tabulated rate >= 0
        bool ok = true;
        for (size_t q = 0; q < (size_t)nprocs * nE && ok; q++) ok = rate[q] >= 0;
        T.v.mono_mask = ok ? ~0ULL : 0ULL;
    }
    rc = upload_procs(ctx, T, procs, nprocs); if (rc) return rc;
    T.smem_bytes = sizeof(ptl_process_desc) * nprocs;
    ctx->tables.push_back(T);
    return (int32_t)ctx->tables.size() - 1;
}

EXPORT int32_t ptl_table_create_linear_vb(ptl_context* ctx, int32_t grid_kind, double L1, double L2, int32_t nE, int32_t nprocs,
                                          const double* rate, const double* ratebound_vec, const ptl_process_desc* procs) {
    PTL_BIND(ctx);
    if (!ctx || !ratebound_vec || nE < 2) return PTL_EINVAL;
    double mx = 0;
    for (int e = 0; e < nE; e++) mx = ratebound_vec[e] > mx ? ratebound_vec[e] : mx;
    int32_t id = ptl_table_create_linear(ctx, grid_kind, L1, L2, nE, nprocs, rate, mx, procs);
    if (id < 0) return id;
    Table& T = ctx->tables[id];
    int32_t rc = upload_doubles(ctx, ratebound_vec, (size_t)nE, &T.d_rbvec); if (rc) return rc;
    T.v.rbvec = T.d_rbvec;
    return id;
}

EXPORT int32_t ptl_cheb_loss_create(ptl_context* ctx, int32_t order, int32_t k, double xmax, const double* ec, const double* pc) {
    PTL_BIND(ctx);
    if (!ctx || order < 1 || order > MAX_ORDER || k < 1) return PTL_EINVAL;
    if (ctx->cls.size() >= (size_t)MAX_CHEBLOSS) return PTL_ENOMEM;
    ChebLoss c;
    c.v.order = order; c.v.k = k; c.v.xmax = xmax; c.v.rxmax = 1.0 / xmax;
    double *de = nullptr, *dp = nullptr;
    int32_t rc = upload_doubles(ctx, ec, (size_t)order * (k + 1), &de); if (rc) return rc;
    rc = upload_doubles(ctx, pc, (size_t)order * (k + 1), &dp); if (rc) return rc;
    c.v.ec = de; c.v.pc = dp;
    ctx->cls.push_back(c);
    return (int32_t)ctx->cls.size() - 1;
}

EXPORT int32_t ptl_table_eval(ptl_context* ctx, int32_t table, int64_t n, const double* energy, double* rates_out, double* bound_out) {
    PTL_BIND(ctx);
    if (!ctx || table < 0 || table >= (int)ctx->tables.size()) return PTL_EHANDLE;
    if (n <= 0) return 0;
    const Table& T = ctx->tables[table];
    size_t np = (size_t)(T.v.nprocs > 0 ? T.v.nprocs : 1);
    int32_t rc = ensure_tmp(ctx, sizeof(double) * (size_t)n * (np + 2)); if (rc) return rc;
    double* d_e = (double*)ctx->d_tmp;
    double* d_b = d_e + n;
    double* d_r = d_b + n;
    CK(cudaMemcpyAsync(d_e, energy, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    k_table_eval<<<grid_for(n, 128), 128, 0, ctx->stream>>>(T.v, n, d_e, d_r, d_b, &ctx->d_sc->flags);
    LAUNCHED();
    if (T.v.nprocs) CK(cudaMemcpyAsync(rates_out, d_r, sizeof(double) * n * T.v.nprocs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(bound_out, d_b, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// =====================================================================================================
// populations
// =====================================================================================================
EXPORT int32_t ptl_population_create(ptl_context* ctx, int32_t species, int64_t capacity, double energy_cut, int32_t table) {
    PTL_BIND(ctx);
    if (!ctx || species < 0 || species >= PTL_NSPECIES || capacity < 1) return PTL_EINVAL;
    if (table < 0 || table >= (int)ctx->tables.size()) return PTL_EHANDLE;
    if (ctx->pops.size() >= 64) return PTL_ENOMEM;
    Pop P;
    size_t colb = align256(sizeof(double) * (size_t)capacity);
    size_t actb = align256((size_t)capacity);
    size_t total = colb * 11 + actb;
    if (cudaMalloc(&P.block, total) != cudaSuccess) { cudaGetLastError(); ctx->err = "cudaMalloc of the particle store failed"; return PTL_ENOMEM; }
    char* base = (char*)P.block;
    for (int c = 0; c < 10; c++) P.v.col[c] = (double*)(base + colb * c);
    P.v.uid = (uint64_t*)(base + colb * 10);
    P.v.active = (uint8_t*)(base + colb * 11);
    P.v.capacity = capacity;
    P.v.energy_cut = energy_cut;
    P.v.species = species;
    P.v.present = 1;
    P.table = table;
    P.slot = (int)ctx->pops.size();
    P.v.n = &ctx->d_sc->pop_n[P.slot];
    P.alive = true;
    CK(cudaMemsetAsync(P.v.n, 0, sizeof(unsigned long long), ctx->stream));
    ctx->pops.push_back(P);
    return (int32_t)ctx->pops.size() - 1;
}

EXPORT int32_t ptl_population_destroy(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(P->block);
    P->block = nullptr;
    P->alive = false;
    return 0;
}

EXPORT int32_t ptl_population_upload(ptl_context* ctx, int32_t pop, int64_t n, const double* x3, const double* p3, const double* w,
                                     const double* t, const double* s, const double* r, const uint8_t* active, const uint64_t* uid) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    if (n < 0 || n > P->v.capacity) return PTL_EINVAL;
    if (n > 0 && (!x3 || !p3 || !w || !t || !s || !r || !active)) return PTL_EINVAL;
    int32_t rc = ensure_stage(ctx, (size_t)n); if (rc) return rc;
    // x, p: host xyz-interleaved -> planar columns through the staging buffers.  When a buffer holds the whole vector (x in
    // stage[0], p in stage[1]) the copies are issued first and the transposes after them: the host arrays are borrowed only
    // until the COPIES are done, and a copy needs no SM — the call does not have to wait for transposes that may be queued
    // behind another context's advance kernel (persistent CTAs that fill every SM).  With several contexts pipelining
    // shards through one GPU this took the upload of a shard from 70 ms (waiting for SMs) to the PCIe time.
    const bool whole = (size_t)n <= ctx->stage_rows;
    if (whole) {
        if (n > 0) {
            CK(cudaMemcpyAsync(ctx->stage[0], x3, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->stage[1], p3, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
        }
    } else {
        int b = 0;
        for (int which = 0; which < 2; which++) {
            const double* src = which == 0 ? x3 : p3;
            int c0 = which == 0 ? COL_X0 : COL_P0;
            for (long long off = 0; off < n; off += (long long)ctx->stage_rows) {
                long long m = n - off < (long long)ctx->stage_rows ? n - off : (long long)ctx->stage_rows;
                CK(cudaMemcpyAsync(ctx->stage[b], src + 3 * off, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, ctx->stream));
                k_aos3_to_planar<<<grid_for(m, 256), 256, 0, ctx->stream>>>((const double*)ctx->stage[b], P->v.col[c0] + off, P->v.col[c0 + 1] + off,
                                                                            P->v.col[c0 + 2] + off, m);
                LAUNCHED();
                b ^= 1;
            }
        }
    }
    bool uid_kernel = false;
    if (n > 0) {
        CK(cudaMemcpyAsync(P->v.col[COL_W], w, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(P->v.col[COL_T], t, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(P->v.col[COL_S], s, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(P->v.col[COL_R], r, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(P->v.active, active, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        if (uid) {   // explicit uids: the default-uid counter must end up past the largest sequential one (folded in lazily)
            CK(cudaMemcpyAsync(P->v.uid, uid, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
            uid_kernel = true;
        } else {
            rc = refresh_next_uid(ctx); if (rc) return rc;
            k_fill_uid<<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v.uid, ctx->next_uid, n);
            LAUNCHED();
            ctx->next_uid += (uint64_t)n;
        }
    }
    // every byte of the host arrays has been read once this event fires
    if (!ctx->ev_copy) CK(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming | cudaEventBlockingSync));
    CK(cudaEventRecord(ctx->ev_copy, ctx->stream));
    if (whole && n > 0) {
        k_aos3_to_planar<<<grid_for(n, 256), 256, 0, ctx->stream>>>((const double*)ctx->stage[0], P->v.col[COL_X0], P->v.col[COL_X1], P->v.col[COL_X2], n);
        LAUNCHED();
        k_aos3_to_planar<<<grid_for(n, 256), 256, 0, ctx->stream>>>((const double*)ctx->stage[1], P->v.col[COL_P0], P->v.col[COL_P1], P->v.col[COL_P2], n);
        LAUNCHED();
    }
    if (uid_kernel) {
        int blocks = (int)((n + 1023) / 1024);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        k_uid_max<<<blocks, 256, 0, ctx->stream>>>(P->v.uid, n, &ctx->d_sc->max_uid);
        LAUNCHED();
    }
    P->iup = 0;
    rc = set_n(ctx, *P, n); if (rc) return rc;
    CK(cudaEventSynchronize(ctx->ev_copy));     // host arrays are only borrowed for the duration of the call
    return 0;
}

EXPORT int64_t ptl_population_download(ptl_context* ctx, int32_t pop, int64_t max_n, double* x3, double* p3, double* w, double* t,
                                       double* s, double* r, uint8_t* active, uint64_t* uid) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (n > max_n) n = max_n;
    if (n <= 0) return 0;
    rc = ensure_stage(ctx); if (rc) return rc;
    int b = 0;
    for (int which = 0; which < 2; which++) {
        double* dst = which == 0 ? x3 : p3;
        if (!dst) continue;
        int c0 = which == 0 ? COL_X0 : COL_P0;
        for (long long off = 0; off < n; off += (long long)ctx->stage_rows) {
            long long m = n - off < (long long)ctx->stage_rows ? n - off : (long long)ctx->stage_rows;
            k_planar_to_aos3<<<grid_for(m, 256), 256, 0, ctx->stream>>>(P->v.col[c0] + off, P->v.col[c0 + 1] + off, P->v.col[c0 + 2] + off,
                                                                        (double*)ctx->stage[b], m);
            LAUNCHED();
            CK(cudaMemcpyAsync(dst + 3 * off, ctx->stage[b], sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, ctx->stream));
            b ^= 1;
        }
    }
    if (w) CK(cudaMemcpyAsync(w, P->v.col[COL_W], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (t) CK(cudaMemcpyAsync(t, P->v.col[COL_T], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (s) CK(cudaMemcpyAsync(s, P->v.col[COL_S], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (r) CK(cudaMemcpyAsync(r, P->v.col[COL_R], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (active) CK(cudaMemcpyAsync(active, P->v.active, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (uid) CK(cudaMemcpyAsync(uid, P->v.uid, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return n;
}

EXPORT int64_t ptl_population_n(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n);
    return rc ? rc : n;
}

EXPORT int64_t ptl_population_capacity(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    return P ? P->v.capacity : PTL_EHANDLE;
}

EXPORT int32_t ptl_population_clear(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    return set_n(ctx, *P, 0);
}

EXPORT int32_t ptl_population_set_n(ptl_context* ctx, int32_t pop, int64_t n) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    if (n < 0 || n > P->v.capacity) return PTL_EINVAL;
    return set_n(ctx, *P, n);
}

EXPORT void* ptl_population_column_ptr(ptl_context* ctx, int32_t pop, int32_t col) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || col < 0 || col >= NCOLS) return nullptr;
    if (col < 10) return P->v.col[col];
    return col == COL_ACTIVE ? (void*)P->v.active : (void*)P->v.uid;
}

EXPORT int64_t ptl_population_append(ptl_context* ctx, int32_t pop, const double* x3, const double* p3, double w, double t, double s,
                                     double r, uint64_t uid) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || !x3 || !p3) return PTL_EHANDLE;
    double p2 = p3[0] * p3[0] + p3[1] * p3[1] + p3[2] * p3[2];
    double eng = P->v.species == PTL_PHOTON ? sqrt(p2) * CO_C
               : P->v.species == PTL_SLOW_ELECTRON ? 0.5 * CO_ME * p2
               : sqrt(CO_MC2 * CO_MC2 + CO_C2 * p2) - CO_MC2;
    if (eng <= P->v.energy_cut) return -1;   // population.jl:105
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (n >= P->v.capacity) {      // @assert n < length(particles)  population.jl:107: sticky bit (OR-ed, other bits survive) + its own code
        k_or_flags<<<1, 1, 0, ctx->stream>>>(&ctx->d_sc->flags, PTL_ERR_CAPACITY_OVERFLOW);
        LAUNCHED();
        ctx->err = "ptl_population_append: population is full";
        return PTL_ECAPACITY;
    }
    double vals[10] = {x3[0], x3[1], x3[2], p3[0], p3[1], p3[2], w, t, s, r};
    for (int c = 0; c < 10; c++) CK(cudaMemcpyAsync(P->v.col[c] + n, &vals[c], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    uint8_t one = 1;
    if (uid == 0) {
        rc = refresh_next_uid(ctx); if (rc) return rc;
        uid = ctx->next_uid++;
    } else if (!(uid & PTL_UID_HASHED_BIT) && uid >= ctx->next_uid) {
        ctx->next_uid = uid + 1;
    }
    CK(cudaMemcpyAsync(P->v.active + n, &one, 1, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(P->v.uid + n, &uid, sizeof(uid), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = set_n(ctx, *P, n + 1); if (rc) return rc;
    return n;
}

EXPORT int32_t ptl_population_deactivate(ptl_context* ctx, int32_t pop, int64_t i) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (i < 0 || i >= n) return PTL_EINVAL;
    CK(cudaMemsetAsync(P->v.active + i, 0, 1, ctx->stream));
    return 0;
}

// ---- droplow! / repack! ---------------------------------------------------------------------------------
namespace {
int64_t compact(ptl_context* ctx, Pop& P, bool do_flag, double thres) {
    long long n = 0;
    int32_t rc = read_n(ctx, P, &n); if (rc) return rc;
    if (n == 0) return 0;
    long long ntiles = (n + CMP_TILE - 1) / CMP_TILE;
    // Scratch is sized by the population's CAPACITY, once: sizing it by the current n re-allocated it almost every step of a
    // growing avalanche, and a cudaFree/cudaMalloc pair inside droplow! cost up to 200 ms (measured: 2 ms -> 200 ms).
    if ((size_t)ntiles > ctx->tiles_cap) {
        cudaFree(ctx->d_tile_counts); cudaFree(ctx->d_tile_offsets);
        size_t cap = (size_t)((P.v.capacity + CMP_TILE - 1) / CMP_TILE) + 1024;
        CK(cudaMalloc(&ctx->d_tile_counts, sizeof(unsigned int) * cap));
        CK(cudaMalloc(&ctx->d_tile_offsets, sizeof(unsigned long long) * cap));
        ctx->tiles_cap = cap;
    }
    double th = thres == 0 ? P.v.energy_cut : thres;   // population.jl:274
    DISPATCH_SPECIES(P.v.species, k_flag_count<SP><<<(unsigned)ntiles, CMP_THREADS, 0, ctx->stream>>>(P.v, n, th, do_flag ? 1 : 0, ctx->d_tile_counts));
    LAUNCHED();
    k_scan_tiles<<<1, 1024, 0, ctx->stream>>>(ctx->d_tile_counts, ntiles, ctx->d_tile_offsets, &ctx->d_sc->total);
    LAUNCHED();
    CK(cudaMemsetAsync(&ctx->d_sc->nmoves, 0, sizeof(unsigned long long), ctx->stream));
    rc = sync_scalars(ctx); if (rc) return rc;
    long long total = (long long)ctx->h_sc->total;
    long long max_moves = total < n - total ? total : n - total;
    if (max_moves > 0) {
        if ((size_t)max_moves > ctx->moves_cap) {
            cudaFree(ctx->d_holes); cudaFree(ctx->d_tails);
            size_t cap = (size_t)(P.v.capacity / 2) + 1024;           // min(actives, dead) <= n / 2 <= capacity / 2
            CK(cudaMalloc(&ctx->d_holes, sizeof(long long) * cap));
            CK(cudaMalloc(&ctx->d_tails, sizeof(long long) * cap));
            ctx->moves_cap = cap;
        }
        k_emit_moves<<<(unsigned)ntiles, CMP_THREADS, 0, ctx->stream>>>(P.v, n, ctx->d_tile_offsets, &ctx->d_sc->total, ctx->d_holes, ctx->d_tails,
                                                                         &ctx->d_sc->nmoves);
        LAUNCHED();
        k_apply_moves<<<grid_for(max_moves, 256), 256, 0, ctx->stream>>>(P.v, ctx->d_holes, ctx->d_tails, &ctx->d_sc->nmoves);
        LAUNCHED();
    }
    rc = set_n(ctx, P, total); if (rc) return rc;
    return total;
}
}  // namespace

EXPORT int64_t ptl_droplow(ptl_context* ctx, int32_t pop, double thres) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    return compact(ctx, *P, true, thres);
}

EXPORT int64_t ptl_repack(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    return compact(ctx, *P, false, 0.0);
}

namespace ptl_host {
int32_t diag_local_launch(ptl_context* ctx, Pop& P, long long n) {
    if (n > 0) {
        int blocks = (int)((n + DIAG_THREADS - 1) / DIAG_THREADS);
        if (blocks > ctx->partial_blocks) blocks = ctx->partial_blocks;
        DISPATCH_SPECIES(P.v.species, k_diag_partial<SP><<<blocks, DIAG_THREADS, 0, ctx->stream>>>(P.v, n, ctx->d_partial));
        LAUNCHED();
        k_diag_final<<<1, 32, 0, ctx->stream>>>(ctx->d_partial, blocks, ctx->d_sc->diag);
        LAUNCHED();
    } else {
        for (int q = 0; q < DIAG_NVAL; q++) ctx->h_sc->diag[q] = 0;
        ctx->h_sc->diag[10] = -INFINITY;
        CK(cudaMemcpyAsync(ctx->d_sc->diag, ctx->h_sc->diag, sizeof(double) * DIAG_NVAL, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->h_sc->diag[11] = (double)n;
    CK(cudaMemcpyAsync(ctx->d_sc->diag + 11, ctx->h_sc->diag + 11, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

void diag_unpack(const double* d, ptl_diag_out* out) {
    memset(out, 0, sizeof(*out));
    out->nactive = (int64_t)llround(d[0]);
    out->weight = d[1]; out->wenergy = d[2];
    for (int c = 0; c < 3; c++) { out->wx[c] = d[3 + c]; out->wx2[c] = d[6 + c]; }
    out->wr2 = d[9];
    out->maxenergy = d[10];
    out->n = (int64_t)llround(d[11]);
}

int32_t histogram_local_launch(ptl_context* ctx, Pop& P, int32_t quantity, double lo, double hi, int32_t nbins, int32_t logscale) {
    long long n = 0;
    int32_t rc = read_n(ctx, P, &n); if (rc) return rc;
    rc = ensure_tmp(ctx, sizeof(double) * nbins); if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_tmp, 0, sizeof(double) * nbins, ctx->stream));
    if (n > 0) {
        int blocks = (int)((n + 255) / 256);
        if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
        DISPATCH_SPECIES(P.v.species, k_histogram<SP><<<blocks, 256, sizeof(double) * nbins, ctx->stream>>>(P.v, n, quantity, lo, hi, nbins, logscale,
                                                                                                            (double*)ctx->d_tmp));
        LAUNCHED();
    }
    return 0;
}
}  // namespace ptl_host

// ---- diagnostics -----------------------------------------------------------------------------------------
EXPORT int32_t ptl_diag(ptl_context* ctx, int32_t pop, ptl_diag_out* out) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || !out) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    rc = diag_local_launch(ctx, *P, n); if (rc) return rc;
    rc = sync_scalars(ctx); if (rc) return rc;
    diag_unpack(ctx->h_sc->diag, out);
    return 0;
}

EXPORT int32_t ptl_histogram(ptl_context* ctx, int32_t pop, int32_t quantity, double lo, double hi, int32_t nbins, int32_t logscale,
                             double* out) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || !out || nbins < 1 || nbins > 4096 || !(hi > lo)) return PTL_EINVAL;
    int32_t rc = histogram_local_launch(ctx, *P, quantity, lo, hi, nbins, logscale); if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->d_tmp, sizeof(double) * nbins, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

namespace {
// stage the nodes of an energy law on the device (n == 1: the constant travels in the struct)
int32_t make_law(ptl_context* ctx, double lo, double hi, int32_t n, int32_t logscale, const double* v, EnergyLaw* L) {
    if (n < 1 || n > (1 << 20) || !v || (n > 1 && !(hi > lo))) return PTL_EINVAL;
    L->lo = lo; L->hi = hi; L->n = n; L->logscale = logscale; L->c = v[0]; L->v = nullptr;
    if (n > 1) {
        int32_t rc = ensure_tmp(ctx, sizeof(double) * (size_t)n); if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->d_tmp, v, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));      // the host array is only borrowed
        L->v = (const double*)ctx->d_tmp;
    }
    return 0;
}
}  // namespace

EXPORT int32_t ptl_roulette_law(ptl_context* ctx, int32_t pop, double lo, double hi, int32_t nn, int32_t logscale, const double* pv) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    EnergyLaw L;
    int32_t rc = make_law(ctx, lo, hi, nn, logscale, pv, &L); if (rc) return rc;
    long long n = 0;
    rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (n > 0) {
        k_roulette<<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v, n, L, ctx->step, (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32));
        LAUNCHED();
    }
    ctx->step++;
    return 0;
}

EXPORT int32_t ptl_roulette(ptl_context* ctx, int32_t pop, double p) {
    if (!(p > 0)) return PTL_EINVAL;
    return ptl_roulette_law(ctx, pop, 0.0, 1.0, 1, 0, &p);
}

EXPORT int32_t ptl_split_law(ptl_context* ctx, int32_t pop, double lo, double hi, int32_t nn, int32_t logscale, const double* pv) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    EnergyLaw L;
    int32_t rc = make_law(ctx, lo, hi, nn, logscale, pv, &L); if (rc) return rc;
    long long n = 0;
    rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (n > 0) {
        k_split<<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v, n, L, ctx->step, (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32), &ctx->d_sc->flags);
        LAUNCHED();
    }
    ctx->step++;
    rc = read_n(ctx, *P, &n); if (rc) return rc;
    return ctx->h_sc->flags;
}

EXPORT int32_t ptl_split(ptl_context* ctx, int32_t pop, double p) {
    if (!(p >= 0)) return PTL_EINVAL;
    return ptl_split_law(ctx, pop, 0.0, 1.0, 1, 0, &p);
}

// shuffle!(popl) population.jl:266-271
EXPORT int32_t ptl_shuffle(ptl_context* ctx, int32_t pop) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (n > 1) {
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long*)nullptr,
                                        (long long*)nullptr, (int)n, 0, 64, ctx->stream);
        const size_t nb = align256(sizeof(unsigned long long) * (size_t)n);
        rc = ensure_tmp(ctx, 5 * nb + align256(sort_bytes)); if (rc) return rc;
        char* base = (char*)ctx->d_tmp;
        unsigned long long *k0 = (unsigned long long*)base, *k1 = (unsigned long long*)(base + nb);
        long long *r0 = (long long*)(base + 2 * nb), *r1 = (long long*)(base + 3 * nb);
        double* col = (double*)(base + 4 * nb);
        void* sort_tmp = base + 5 * nb;
        k_shuffle_keys<<<grid_for(n, 256), 256, 0, ctx->stream>>>(k0, r0, n, ctx->step, (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32));
        LAUNCHED();
        // stable LSD radix sort: equal keys keep their row order, like the oracle's (key, row) comparison
        CK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, k0, k1, r0, r1, (int)n, 0, 64, ctx->stream));
        ctx->launch_total++;
        for (int c = 0; c < 10; c++) {
            k_gather<double><<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v.col[c], r1, col, n);
            LAUNCHED();
            CK(cudaMemcpyAsync(P->v.col[c], col, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        k_gather<uint64_t><<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v.uid, r1, (uint64_t*)col, n);
        LAUNCHED();
        CK(cudaMemcpyAsync(P->v.uid, col, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        k_gather<uint8_t><<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->v.active, r1, (uint8_t*)col, n);
        LAUNCHED();
        CK(cudaMemcpyAsync(P->v.active, col, (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->step++;
    return 0;
}

// =====================================================================================================
// multi-population + advance
// =====================================================================================================
EXPORT int32_t ptl_multipop_create(ptl_context* ctx, const int32_t* pops, int32_t count) {
    PTL_BIND(ctx);
    if (!ctx || !pops || count < 1 || count > PTL_NSPECIES) return PTL_EINVAL;
    MultiPop m;
    for (int s = 0; s < PTL_NSPECIES; s++) m.by_species[s] = -1;
    for (int i = 0; i < count; i++) {
        Pop* P = get_pop(ctx, pops[i]);
        if (!P) return PTL_EHANDLE;
        if (m.by_species[P->v.species] >= 0) { ctx->err = "two populations of the same species in one MultiPopulation"; return PTL_EINVAL; }
        m.by_species[P->v.species] = pops[i];
        m.pops.push_back(pops[i]);
    }
    ctx->mps.push_back(m);
    return (int32_t)ctx->mps.size() - 1;
}

EXPORT int32_t ptl_init(ptl_context* ctx, int32_t mp) {
    PTL_BIND(ctx);
    if (!ctx || mp < 0 || mp >= (int)ctx->mps.size()) return PTL_EHANDLE;
    const MultiPop& M = ctx->mps[mp];
    AdvanceParams A;
    fill_params(ctx, &M, A);
    int32_t rc = sync_scalars(ctx); if (rc) return rc;
    for (int pi : M.pops) {
        Pop& P = ctx->pops[pi];
        long long n = (long long)ctx->h_sc->pop_n[P.slot];
        if (n <= 0) continue;
        DISPATCH_SPECIES(P.v.species, k_init_r<SP><<<grid_for(n, 256), 256, 0, ctx->stream>>>(A, n));
        LAUNCHED();
    }
    rc = sync_scalars(ctx); if (rc) return rc;
    return ctx->h_sc->flags;
}

namespace {
int32_t ensure_wall(ptl_context* ctx, int k, long long capacity) {
    Wall& W = ctx->walls[k];
    if (W.block && W.b.capacity >= capacity) return 0;
    if (W.block) cudaFree(W.block);
    size_t colb = align256(sizeof(double) * (size_t)capacity);
    CK(cudaMalloc(&W.block, colb * 8));
    for (int c = 0; c < 8; c++) W.b.col[c] = (double*)((char*)W.block + colb * c);
    W.b.capacity = capacity;
    W.b.n = &ctx->d_sc->wall_n[k];
    CK(cudaMemsetAsync(W.b.n, 0, sizeof(unsigned long long), ctx->stream));
    return 0;
}
}  // namespace

EXPORT int32_t ptl_advance(ptl_context* ctx, int32_t mp, const ptl_pusher_desc* pusher, double tfinal, const ptl_callback_desc* cb) {
    PTL_BIND(ctx);
    if (!ctx || mp < 0 || mp >= (int)ctx->mps.size() || !pusher) return PTL_EHANDLE;
    if (pusher->nforcings < 0 || pusher->nforcings > PTL_MAX_FORCINGS) return PTL_EINVAL;
    const MultiPop& M = ctx->mps[mp];
    bool has_cb = cb && (cb->nwalls > 0 || cb->count_collisions);
    if (cb && (cb->nwalls < 0 || cb->nwalls > PTL_MAX_WALLS)) return PTL_EINVAL;
    // descriptors come from a foreign caller: everything the kernels index with is range-checked here
    if (pusher->kind != PTL_PUSHER_NULL && pusher->kind != PTL_PUSHER_RK2) { ctx->err = "unknown pusher kind"; return PTL_EINVAL; }
    for (int k = 0; k < pusher->nforcings; k++) {
        const ptl_forcing_desc& f = pusher->forcing[k];
        if (f.kind < PTL_FORCE_NONE || f.kind > PTL_FORCE_CHEB_CONTINUUM) { ctx->err = "unknown forcing kind"; return PTL_EINVAL; }
        if (f.kind == PTL_FORCE_EM && (f.e.kind < PTL_FIELD_ZERO || f.e.kind > PTL_FIELD_CONFINED_DL || f.b.kind < PTL_FIELD_ZERO || f.b.kind > PTL_FIELD_CONFINED_DL)) {
            ctx->err = "unknown field kind";
            return PTL_EINVAL;
        }
        if (f.kind == PTL_FORCE_CHEB_CONTINUUM && (f.cheb_id < 0 || f.cheb_id >= (int)ctx->cls.size())) {
            ctx->err = "forcing refers to a ChebContinuumLoss id that was never created";
            return PTL_EHANDLE;
        }
    }
    if (has_cb) {
        for (int k = 0; k < cb->nwalls; k++) {
            if (cb->wall[k].coord < 0 || cb->wall[k].coord > 2 || cb->wall[k].species < 0 || cb->wall[k].species >= PTL_NSPECIES) {
                ctx->err = "wall callback: coord must be 0..2 and species a valid species id";
                return PTL_EINVAL;
            }
        }
    }
    if (has_cb) {
        for (int k = 0; k < cb->nwalls; k++) {
            int pi = (cb->wall[k].species >= 0 && cb->wall[k].species < PTL_NSPECIES) ? M.by_species[cb->wall[k].species] : -1;
            long long cap = pi >= 0 ? ctx->pops[pi].v.capacity : 1024;
            if (cap > (1LL << 22)) cap = 1LL << 22;
            int32_t rc = ensure_wall(ctx, k, cap); if (rc) return rc;
        }
    }
    AdvanceParams A;
    fill_params(ctx, &M, A);
    A.pusher = *pusher;
    if (has_cb) A.cb = *cb;
    A.has_cb = has_cb ? 1 : 0;
    A.tfinal = tfinal;
    // canonicalise: a HomogeneousField of zeros is no field (e + v x 0 == e), and a lone uniform E is the hot path
    for (int k = 0; k < A.pusher.nforcings; k++) {
        ptl_forcing_desc& f = A.pusher.forcing[k];
        if (f.kind == PTL_FORCE_EM && f.b.kind == PTL_FIELD_HOMOGENEOUS && f.b.par[0] == 0 && f.b.par[1] == 0 && f.b.par[2] == 0)
            f.b.kind = PTL_FIELD_ZERO;
    }
    A.fast_force = 0;
    if (A.pusher.kind == PTL_PUSHER_RK2 && A.pusher.nforcings == 1 && A.pusher.forcing[0].kind == PTL_FORCE_EM &&
        A.pusher.forcing[0].species_mask == 0 && A.pusher.forcing[0].e.kind == PTL_FIELD_HOMOGENEOUS &&
        A.pusher.forcing[0].b.kind == PTL_FIELD_ZERO) {
        A.fast_force = 1;
        for (int c = 0; c < 3; c++) A.fastE[c] = A.pusher.forcing[0].e.par[c] * CO_E;
    }
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    CK(cudaMemsetAsync(ctx->d_sc->substeps, 0, (PTL_NSPECIES + 1) * sizeof(unsigned long long), ctx->stream));   // substeps[], births
    for (int pi : M.pops) { ctx->pops[pi].iup = 0; ctx->pops[pi].rows_last = 0; }   // advance_init!: iup = 1  (mixed_population.jl:101)
    bool first = true;
    static const bool trace = getenv("PTL_TRACE") != nullptr;   // per-pass wall time / rows / sub-steps on stderr
    auto tprev = std::chrono::steady_clock::now();
    long long rows_prev = 0; unsigned long long sub_prev = 0;
    for (;;) {
        int32_t rc = sync_scalars(ctx); if (rc) return rc;     // read every popl.n (one host sync per pass)
        if (trace) {
            auto tnow = std::chrono::steady_clock::now();
            unsigned long long sub = 0;
            for (int sp = 0; sp < PTL_NSPECIES; sp++) sub += ctx->h_sc->substeps[sp];
            if (ctx->stats.passes > 0)
                fprintf(stderr, "[ptl trace] step %llu pass %d: rows %lld substeps %llu  %.3f ms\n", (unsigned long long)ctx->step,
                        (int)ctx->stats.passes, rows_prev, sub - sub_prev, std::chrono::duration<double, std::milli>(tnow - tprev).count());
            if (ctx->stats.passes > 0 && ctx->h_sc->dbg[2])
                fprintf(stderr, "[ptl trace]     rounds per CTA: max %llu mean %.0f over %llu CTAs\n", ctx->h_sc->dbg[0],
                        (double)ctx->h_sc->dbg[1] / (double)ctx->h_sc->dbg[2], ctx->h_sc->dbg[2]);
            if (ctx->stats.passes > 0 && ctx->h_sc->dbg[2]) {
                const unsigned long long* g = ctx->h_sc->dbg;
                double v[5]; memcpy(v, g + 10, sizeof(v));
                fprintf(stderr, "[ptl trace]     lonely rounds by class: %llu %llu %llu %llu %llu %llu; last lonely row %llu p=(%.4e %.4e %.4e) r=%.4e s=%.4e\n",
                        g[4], g[5], g[6], g[7], g[8], g[9], g[3], v[0], v[1], v[2], v[3], v[4]);
            }
#ifdef BQ_PROFILE
            if (ctx->stats.passes > 0 && ctx->h_sc->dbg[2]) {   // per-class chunk timing (clock64) of the list-scheduled kernel
                const unsigned long long* g = ctx->h_sc->dbg;
                static const char* nm[6] = {"LOAD", "STEP", "COULOMB", "RBEB", "IONFIN", "OTHER"};
                for (int c = 0; c < 6; c++)
                    if (g[24 + c]) fprintf(stderr, "[ptl trace]     %-8s chunks %10llu  mean %7.0f cyc  max %8llu cyc  share %5.1f %%\n", nm[c], g[24 + c],
                                           (double)g[16 + c] / (double)g[24 + c], g[32 + c], 100.0 * (double)g[16 + c] / (double)(g[42] ? g[42] : 1));
                fprintf(stderr, "[ptl trace]     rounds %llu: warp-cycles total %.3e, in units %.1f %%, at the barrier %.1f %%, max unit per round mean %.0f cyc, round mean %.0f cyc\n",
                        g[41] / 8, (double)g[42], 100.0 * (double)(g[16] + g[17] + g[18] + g[19] + g[20] + g[21]) / (double)(g[42] ? g[42] : 1),
                        100.0 * (double)g[40] / (double)(g[42] ? g[42] : 1), (double)g[43] / (double)(g[41] ? g[41] : 1) , (double)g[42] / (double)(g[41] ? g[41] : 1));
            }
#endif
            CK(cudaMemsetAsync(ctx->d_sc->dbg, 0, sizeof(ctx->d_sc->dbg), ctx->stream));
            tprev = tnow; sub_prev = sub;
        }
        long long total = 0;
        // The species of one pass are independent of each other: a kernel visits rows [iup, n) of its own population and
        // births only append beyond n.  So the kernels of a pass run concurrently — the first species on the context's
        // stream, the others on auxiliary streams forked from it and joined back before the counters are read.  In the
        // latency regime (a few thousand rows, bound by the collision chain of the slowest particle) the chains of the
        // species overlap instead of adding up.  (Sequential when a population overflowed: its count is being clamped.)
        bool clamp = false;
        int nlaunch = 0;
        for (int pi : M.pops) {
            Pop& P = ctx->pops[pi];
            long long n = (long long)ctx->h_sc->pop_n[P.slot];
            if (n > P.v.capacity) clamp = true;
            if ((n > P.v.capacity ? P.v.capacity : n) - P.iup > 0) nlaunch++;
        }
        const bool overlap = ctx->overlap_species && !clamp && nlaunch > 1;
        if (overlap) {
            if (!ctx->ev_fork) CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        }
        int launched = 0;
        for (int pi : M.pops) {
            Pop& P = ctx->pops[pi];
            long long n = (long long)ctx->h_sc->pop_n[P.slot];
            if (n > P.v.capacity) {   // children beyond capacity were dropped and flagged by the kernel
                n = P.v.capacity;
                rc = set_n(ctx, P, n); if (rc) return rc;
            }
            long long rows = n - P.iup;
            if (rows > 0) {
                const Table& T = ctx->tables[P.table];
                const int sp = P.v.species;
                ctx->lstream = ctx->stream;
                if (overlap && launched > 0) {
                    if (!ctx->aux[sp]) {
                        CK(cudaStreamCreateWithFlags(&ctx->aux[sp], cudaStreamNonBlocking));
                        CK(cudaEventCreateWithFlags(&ctx->ev_join[sp], cudaEventDisableTiming));
                    }
                    ctx->lstream = ctx->aux[sp];
                    CK(cudaStreamWaitEvent(ctx->lstream, ctx->ev_fork, 0));
                }
                rc = launch_advance(ctx, sp, A, P.iup, n, first, has_cb, T.smem_bytes, first && P.kappa_est >= 0 && P.kappa_est < 2.0);
                if (ctx->lstream != ctx->stream) {
                    CK(cudaEventRecord(ctx->ev_join[sp], ctx->lstream));
                    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[sp], 0));
                }
                ctx->lstream = ctx->stream;
                if (rc) return rc;
                launched++;
                total += rows;
                ctx->stats.rows += rows;
                P.rows_last += rows;
                P.iup = n;
            }
        }
        ctx->stats.passes++;
        rows_prev = total;
        first = false;
        if (total == 0) break;     // advance1! returned 0 (mixed_population.jl:44-46)
    }
    ctx->step++;
    if (ctx->ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.main_ms = ms;
        ctx->ev_pending = false;
    }
    ctx->stats.substeps = 0;
    for (int sp = 0; sp < PTL_NSPECIES; sp++) ctx->stats.substeps += (int64_t)ctx->h_sc->substeps[sp];
    // sub-steps per row of each population: picks the kernel of its next first pass (streaming below ~2, wavefront above)
    for (int pi : M.pops) {
        Pop& P = ctx->pops[pi];
        if (P.rows_last > 0) P.kappa_est = (double)ctx->h_sc->substeps[P.v.species] / (double)P.rows_last;
    }
    ctx->stats.births = (int64_t)ctx->h_sc->births;
    return ctx->h_sc->flags;
}

EXPORT int32_t ptl_set_profiling(ptl_context* ctx, int32_t on) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    if (on && !ctx->ev0) {
        CK(cudaEventCreate(&ctx->ev0));
        CK(cudaEventCreate(&ctx->ev1));
    }
    ctx->profiling = on != 0;
    return 0;
}

EXPORT int64_t ptl_launch_count(ptl_context* ctx, int32_t reset) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    long long v = ctx->launch_total;
    if (reset) ctx->launch_total = 0;
    return v;
}

EXPORT int32_t ptl_last_advance_stats(ptl_context* ctx, ptl_advance_stats* out) {
    PTL_BIND(ctx);
    if (!ctx || !out) return PTL_EINVAL;
    *out = ctx->stats;
    return 0;
}

EXPORT int32_t ptl_collision_counts(ptl_context* ctx, int32_t table, int64_t* counts, int32_t clear) {
    PTL_BIND(ctx);
    if (!ctx || table < 0 || table >= (int)ctx->tables.size() || !counts) return PTL_EHANDLE;
    Table& T = ctx->tables[table];
    size_t bytes = sizeof(unsigned long long) * (T.v.nprocs + 1);
    CK(cudaMemcpyAsync(counts, T.d_counts, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (clear) CK(cudaMemsetAsync(T.d_counts, 0, bytes, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

EXPORT int64_t ptl_wall_records(ptl_context* ctx, int32_t iwall, int64_t max_n, double* x3, double* p3, double* w, double* t, int32_t clear) {
    PTL_BIND(ctx);
    if (!ctx || iwall < 0 || iwall >= PTL_MAX_WALLS) return PTL_EINVAL;
    Wall& W = ctx->walls[iwall];
    if (!W.block) return 0;
    int32_t rc = sync_scalars(ctx); if (rc) return rc;
    long long total = (long long)ctx->h_sc->wall_n[iwall];
    if (total > W.b.capacity) total = W.b.capacity;
    long long n = total < max_n ? total : max_n;
    if (n > 0) {
        rc = ensure_tmp(ctx, sizeof(double) * 3 * (size_t)n); if (rc) return rc;
        for (int which = 0; which < 2; which++) {
            double* dst = which == 0 ? x3 : p3;
            if (!dst) continue;
            int c0 = which * 3;
            k_planar_to_aos3<<<grid_for(n, 256), 256, 0, ctx->stream>>>(W.b.col[c0], W.b.col[c0 + 1], W.b.col[c0 + 2], (double*)ctx->d_tmp, n);
            LAUNCHED();
            CK(cudaMemcpyAsync(dst, ctx->d_tmp, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (w) CK(cudaMemcpyAsync(w, W.b.col[6], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (t) CK(cudaMemcpyAsync(t, W.b.col[7], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (clear) CK(cudaMemsetAsync(W.b.n, 0, sizeof(unsigned long long), ctx->stream));
    return total;
}

// =====================================================================================================
// test / diagnostic entry points
// =====================================================================================================
EXPORT int32_t ptl_collide_test(ptl_context* ctx, int32_t species, int32_t table, int32_t j, int64_t n, const double* p3, uint64_t uid0,
                                double* out) {
    PTL_BIND(ctx);
    if (!ctx || table < 0 || table >= (int)ctx->tables.size()) return PTL_EHANDLE;
    const Table& T = ctx->tables[table];
    if (j < 0 || j >= T.v.nprocs || species < 0 || species >= PTL_NSPECIES || n < 0) return PTL_EINVAL;
    if (n == 0) return 0;
    AdvanceParams A;
    fill_params(ctx, nullptr, A);
    int32_t rc = ensure_tmp(ctx, sizeof(double) * (size_t)n * 27); if (rc) return rc;
    double* d_p = (double*)ctx->d_tmp;
    double* d_o = d_p + 3 * n;
    CK(cudaMemcpyAsync(d_p, p3, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    DISPATCH_SPECIES(species, k_collide_test<SP><<<grid_for(n, 128), 128, 0, ctx->stream>>>(A, T.v, j, n, d_p, uid0, d_o));
    LAUNCHED();
    CK(cudaMemcpyAsync(out, d_o, sizeof(double) * 24 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

namespace {
__global__ void k_rng_test(unsigned long long uid, uint32_t step, uint32_t seed_lo, uint32_t seed_hi, int n, double* out) {
    Rng rng;
    rng.init(uid, DOM_COLLISION);
    for (int i = 0; i < n; i++) out[i] = rng.u(step, seed_lo, seed_hi);
}
}  // namespace

EXPORT int32_t ptl_rng_test(ptl_context* ctx, uint64_t uid, uint64_t seed, uint32_t step, int32_t n, double* out) {
    PTL_BIND(ctx);
    if (!ctx || n < 0 || !out) return PTL_EINVAL;
    if (n == 0) return 0;
    int32_t rc = ensure_tmp(ctx, sizeof(double) * n); if (rc) return rc;
    k_rng_test<<<1, 1, 0, ctx->stream>>>(uid, step, (uint32_t)seed, (uint32_t)(seed >> 32), n, (double*)ctx->d_tmp);
    LAUNCHED();
    CK(cudaMemcpyAsync(out, ctx->d_tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
