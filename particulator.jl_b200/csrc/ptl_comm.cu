// ptl_comm.cu — the multi-GPU entry points of the C ABI (SURVEY.md section 8b "threading", 8e): one context per GPU plus a
// communicator.  Particles never interact, so the advance path has no data-path collective; NCCL is used for exactly the
// two exchange steps the path has:
//   * reducing the diagnostics the reference prints / decides on every output step (global counts, weights, energy and
//     position moments, spectra: src/run.jl:31-40, src/callback.jl:203,217,239,263) — ptl_diag_allreduce,
//     ptl_histogram_allreduce, ptl_comm_allreduce_f64;
//   * periodic population rebalancing — ptl_rebalance: all-gather of the per-rank counts, a deterministic plan computed
//     identically on every rank, ONE grouped ncclSend/ncclRecv over the 12 column tails, straight out of / into the
//     device-resident SoA columns.
// NCCL is bound at run time (dlopen of libnccl.so.2, or $PTL_NCCL_LIB): a process that already loaded it (torch) shares
// that copy, and the single-GPU path never needs it.  A host in any language drives this with a 128-byte unique id
// (ptl_comm_unique_id on rank 0, shipped to the other ranks by whatever the host has) — no torch types anywhere.
#include <dlfcn.h>
#include <nccl.h>

#include "ptl_host.h"

using namespace ptl;
using namespace ptl_host;

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {getenv("PTL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.err = "libnccl.so.2 not found (set PTL_NCCL_LIB)"; return nullptr; }
#define BIND(field, sym) do { *(void**)(&api.field) = dlsym(api.handle, sym); if (!api.field) { api.err = std::string("missing symbol ") + sym; dlclose(api.handle); api.handle = nullptr; return nullptr; } } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce"); BIND(AllGather, "ncclAllGather"); BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    return &api;
}

bool nccl_ok(ptl_context* ctx, NcclApi* api, ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    ctx->err = std::string(what) + ": " + api->GetErrorString(r);
    return false;
}
#define NK(call) do { if (!nccl_ok(ctx, api, (api->call), #call)) return PTL_ECOMM; } while (0)

static_assert(sizeof(ncclUniqueId) == PTL_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

}  // namespace

// ---- the deterministic plan (pure host code: callable without a GPU, unit-tested on CPU) -----------------------------------
// counts[r] = particles on rank r.  moves_out[3*m] = (src, dst, k): move the LAST k rows of src to the end of dst.  Greedy
// matching of surpluses to deficits in rank order; nothing moves while every rank is within `tolerance` of the mean.
EXPORT int32_t ptl_rebalance_plan(const int64_t* counts, int32_t nranks, double tolerance, int64_t* moves_out, int32_t max_moves) {
    if (!counts || nranks < 1 || (!moves_out && max_moves > 0)) return PTL_EINVAL;
    long long total = 0;
    for (int r = 0; r < nranks; r++) { if (counts[r] < 0) return PTL_EINVAL; total += counts[r]; }
    const long long base = total / nranks, extra = total % nranks;
    const double mean = (double)total / nranks;
    double dev = 0;
    for (int r = 0; r < nranks; r++) dev = fmax(dev, fabs((double)counts[r] - mean));
    if (mean == 0 || dev <= tolerance * mean) return 0;
    std::vector<long long> diff(nranks);
    for (int r = 0; r < nranks; r++) diff[r] = counts[r] - (base + (r < extra ? 1 : 0));
    int i = 0, j = 0, m = 0;
    while (true) {
        while (i < nranks && diff[i] <= 0) i++;
        while (j < nranks && diff[j] >= 0) j++;
        if (i >= nranks || j >= nranks) break;
        long long k = diff[i] < -diff[j] ? diff[i] : -diff[j];
        if (m >= max_moves) return PTL_EINVAL;
        moves_out[3 * m] = i; moves_out[3 * m + 1] = j; moves_out[3 * m + 2] = k;
        m++;
        diff[i] -= k; diff[j] += k;
    }
    return m;
}

EXPORT int32_t ptl_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return PTL_EINVAL;
    NcclApi* api = nccl_api();
    if (!api) return PTL_ECOMM;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return PTL_ECOMM;
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

EXPORT int32_t ptl_comm_init(ptl_context* ctx, const uint8_t* id, int32_t rank, int32_t nranks) {
    PTL_BIND(ctx);
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return PTL_EINVAL;
    if (ctx->comm) { ctx->err = "communicator already initialised"; return PTL_EINVAL; }
    NcclApi* api = nccl_api();
    if (!api) { ctx->err = "NCCL runtime not available"; return PTL_ECOMM; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    NK(CommInitRank(&comm, nranks, uid, rank));
    ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
    if (!ctx->d_coll) CK(cudaMalloc(&ctx->d_coll, sizeof(double) * PTL_COLL_SCRATCH));
    // NCCL sets up point-to-point connections lazily, at the first send / recv between a pair (tens of ms each): a
    // rebalance that pairs two ranks for the first time would pay that inside a time step.  Connect every pair now with
    // one grouped exchange of a single element per peer (measured on 8 B200: 43 ms -> see DESIGN.md section 7 per step).
    if (nranks > 1 && nranks <= PTL_COLL_SCRATCH / 2) {
        NK(GroupStart());
        for (int peer = 0; peer < nranks; peer++) {
            if (peer == rank) continue;
            NK(Send(ctx->d_coll + peer, 1, ncclDouble, peer, comm, ctx->stream));
            NK(Recv(ctx->d_coll + nranks + peer, 1, ncclDouble, peer, comm, ctx->stream));
        }
        NK(GroupEnd());
        NK(AllReduce(ctx->d_coll, ctx->d_coll, 1, ncclDouble, ncclSum, comm, ctx->stream));      // and the collective channels
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // default uids of this context never collide with another rank's (every context used to start at 1)
    if (ctx->next_uid < ((uint64_t)rank << 40) + 1) ctx->next_uid = ((uint64_t)rank << 40) + 1;
    return 0;
}

EXPORT int32_t ptl_comm_destroy(ptl_context* ctx) {
    PTL_BIND(ctx);
    if (!ctx) return PTL_EINVAL;
    if (ctx->comm) {
        NcclApi* api = nccl_api();
        cudaStreamSynchronize(ctx->stream);
        if (api) api->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->rank = 0; ctx->nranks = 1;
    return 0;
}

EXPORT int32_t ptl_comm_info(ptl_context* ctx, int32_t* rank, int32_t* nranks) {
    if (!ctx) return PTL_EINVAL;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return ctx->comm ? 1 : 0;
}

// in-place all-reduce of a small host vector (op 0 = sum, 1 = max, 2 = min): the global counts RouletteCallback /
// PopulationTargetCallback decide on (src/callback.jl:217,239,263).  Without a communicator it is the identity.
EXPORT int32_t ptl_comm_allreduce_f64(ptl_context* ctx, double* inout, int32_t n, int32_t op) {
    PTL_BIND(ctx);
    if (!ctx || !inout || n < 0 || n > PTL_COLL_SCRATCH || op < 0 || op > 2) return PTL_EINVAL;
    if (!ctx->comm || n == 0) return 0;
    NcclApi* api = nccl_api();
    CK(cudaMemcpyAsync(ctx->d_coll, inout, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    NK(AllReduce(ctx->d_coll, ctx->d_coll, n, ncclDouble, op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin), (ncclComm_t)ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(inout, ctx->d_coll, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ptl_diag with GLOBAL values: the local fused reduction leaves its 12-vector on the device; sums and the max are reduced
// in place by one NCCL group (run.jl:31-40 prints nparticles / nactives / centroids of the whole swarm).
EXPORT int32_t ptl_diag_allreduce(ptl_context* ctx, int32_t pop, ptl_diag_out* out) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || !out) return PTL_EHANDLE;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    rc = diag_local_launch(ctx, *P, n); if (rc) return rc;          // leaves d_sc->diag[0..10]; [11] = n
    if (ctx->comm) {
        NcclApi* api = nccl_api();
        double* d = ctx->d_sc->diag;
        NK(GroupStart());
        NK(AllReduce(d, d, 10, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
        NK(AllReduce(d + 10, d + 10, 1, ncclDouble, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
        NK(AllReduce(d + 11, d + 11, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
        NK(GroupEnd());
        ctx->launch_total += 3;
    }
    rc = sync_scalars(ctx); if (rc) return rc;
    diag_unpack(ctx->h_sc->diag, out);
    return 0;
}

EXPORT int32_t ptl_histogram_allreduce(ptl_context* ctx, int32_t pop, int32_t quantity, double lo, double hi, int32_t nbins, int32_t logscale,
                                       double* out) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P || !out || nbins < 1 || nbins > 4096 || !(hi > lo)) return PTL_EINVAL;
    int32_t rc = histogram_local_launch(ctx, *P, quantity, lo, hi, nbins, logscale); if (rc) return rc;   // bins in d_tmp
    if (ctx->comm) {
        NcclApi* api = nccl_api();
        NK(AllReduce(ctx->d_tmp, ctx->d_tmp, nbins, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
        ctx->launch_total++;
    }
    CK(cudaMemcpyAsync(out, ctx->d_tmp, sizeof(double) * nbins, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Rebalance one species across the ranks of the communicator.  Returns the local n after the exchange (>= 0) or a
// negative error.  `moved_out` (may be NULL) receives the rows this rank sent (+) or received (-).
EXPORT int64_t ptl_rebalance(ptl_context* ctx, int32_t pop, double tolerance, int64_t* moved_out) {
    PTL_BIND(ctx);
    Pop* P = get_pop(ctx, pop);
    if (!P) return PTL_EHANDLE;
    if (moved_out) *moved_out = 0;
    long long n = 0;
    int32_t rc = read_n(ctx, *P, &n); if (rc) return rc;
    if (!ctx->comm || ctx->nranks == 1) return n;
    if (ctx->nranks > PTL_COLL_SCRATCH / 2) return PTL_EINVAL;
    NcclApi* api = nccl_api();
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    // counts of every rank (int64 through the double scratch: bit copies, no arithmetic)
    long long* d_counts = reinterpret_cast<long long*>(ctx->d_coll);
    CK(cudaMemcpyAsync(d_counts + ctx->nranks + ctx->rank, &n, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    NK(AllGather(d_counts + ctx->nranks + ctx->rank, d_counts, 1, ncclInt64, comm, ctx->stream));
    ctx->launch_total++;
    std::vector<int64_t> counts(ctx->nranks);
    CK(cudaMemcpyAsync(counts.data(), d_counts, sizeof(long long) * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> moves(3 * (size_t)ctx->nranks);
    int32_t m = ptl_rebalance_plan(counts.data(), ctx->nranks, tolerance, moves.data(), ctx->nranks);
    if (m < 0) return m;
    long long incoming = 0;
    for (int q = 0; q < m; q++) if (moves[3 * q + 1] == ctx->rank) incoming += moves[3 * q + 2];
    // every rank evaluates the same capacity test on the same plan, so either all of them exchange or none does
    // (capacities are the caller's: equal on all ranks in every use here; a rank that cannot take its share is an error)
    if (n + incoming > P->v.capacity) { ctx->err = "rebalance: receiving rank lacks capacity"; return PTL_ENOMEM; }
    long long sent = 0, received = 0;
    if (m > 0) {
        NK(GroupStart());
        for (int q = 0; q < m; q++) {
            const int src = (int)moves[3 * q], dst = (int)moves[3 * q + 1];
            const long long k = moves[3 * q + 2];
            if (src == ctx->rank) {
                const long long r0 = n - sent - k;                         // the last k rows not yet given away
                for (int c = 0; c < 10; c++) NK(Send(P->v.col[c] + r0, (size_t)k, ncclDouble, dst, comm, ctx->stream));
                NK(Send(P->v.active + r0, (size_t)k, ncclUint8, dst, comm, ctx->stream));
                NK(Send(P->v.uid + r0, (size_t)k, ncclUint64, dst, comm, ctx->stream));
                sent += k;
            } else if (dst == ctx->rank) {
                const long long r0 = n + received;
                for (int c = 0; c < 10; c++) NK(Recv(P->v.col[c] + r0, (size_t)k, ncclDouble, src, comm, ctx->stream));
                NK(Recv(P->v.active + r0, (size_t)k, ncclUint8, src, comm, ctx->stream));
                NK(Recv(P->v.uid + r0, (size_t)k, ncclUint64, src, comm, ctx->stream));
                received += k;
            }
        }
        NK(GroupEnd());
        ctx->launch_total++;
    }
    const long long n_new = n - sent + received;
    if (n_new != n) { rc = set_n(ctx, *P, n_new); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    if (moved_out) *moved_out = sent - received;
    return n_new;
}
