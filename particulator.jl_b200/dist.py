"""Multi-GPU plumbing: one process per GPU, particles sharded, NCCL only where the path has an exchange step.

The reference has no distributed code at all; particles never interact, so every rank
owns an independent shard of every species and advances it with no data-path collective.  Two collectives exist, and
both live INSIDE the shared library (csrc/ptl_comm.cu, NCCL called from the C ABI) so that a Julia host can use them:
  * `diag_allreduce` / `histogram_allreduce` / `allreduce` — sum / max of the fused diagnostics (counts, weights, energy
    and position moments, spectra) so that `nactives`, `meanenergy`, `spread`, ... report GLOBAL values
    (run.jl:31-40, callback.jl:203);
  * `rebalance` — periodic population rebalancing: all-gather of the per-rank counts, a deterministic transfer plan,
    ONE grouped ncclSend/ncclRecv of the 12 column tails straight out of / into the device-resident columns.
This module is the thin host-side caller: `init_comm` ships the 128-byte NCCL id between ranks (through
`torch.distributed` when that is the launcher, or any callable), the rest are one-line calls into the ABI.
`plan_rebalance` / `exchange_columns` restate the host logic in Python over a `torch.distributed` group; with the `gloo`
backend and CPU tensors they are the unit-testable mirror of what ptl_rebalance does (tests/test_dist_gloo.py)."""
import ctypes as C

import numpy as np

from ._lib import DiagOut, dptr


def init_comm(ctx, dist=None, rank=None, nranks=None, exchange=None):
    """Attach an NCCL communicator to `ctx` (ptl_comm_init).  The id is created on rank 0 (ptl_comm_unique_id) and shipped
    with `exchange(bytes_or_None) -> bytes` if given, else broadcast through the torch.distributed group `dist`."""
    import numpy as _np
    if dist is not None:
        rank = dist.get_rank() if rank is None else rank
        nranks = dist.get_world_size() if nranks is None else nranks
    buf = _np.zeros(128, dtype=_np.uint8)
    if rank == 0:
        ctx.check(ctx.backend.comm_unique_id(buf.ctypes.data_as(C.POINTER(C.c_uint8))), "comm_unique_id")
    if exchange is not None:
        buf = _np.frombuffer(exchange(buf.tobytes() if rank == 0 else None), dtype=_np.uint8).copy()
    elif dist is not None and nranks > 1:
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(buf).to(dev)
        dist.broadcast(t, src=0)
        buf = t.cpu().numpy().copy()
    ctx.check(ctx.backend.comm_init(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_uint8)), int(rank), int(nranks)), "comm_init")
    return rank, nranks


def destroy_comm(ctx):
    ctx.check(ctx.backend.comm_destroy(ctx.h), "comm_destroy")


def diag_allreduce(popl):
    """ptl_diag with global sums / max over the communicator of the population's context."""
    d = DiagOut()
    popl.ctx.check(popl.ctx.backend.diag_allreduce(popl.ctx.h, popl.id, C.byref(d)), "diag_allreduce")
    return d


def histogram_allreduce(popl, quantity, lo, hi, nbins, logscale=False):
    out = np.zeros(nbins)
    q = {"energy": 0, "costheta": 1}[quantity]
    popl.ctx.check(popl.ctx.backend.histogram_allreduce(popl.ctx.h, popl.id, q, float(lo), float(hi), nbins, 1 if logscale else 0,
                                                        dptr(out)), "histogram_allreduce")
    return out


def allreduce(ctx, values, op="sum"):
    """In-place all-reduce of a small host vector through the library's communicator."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    ctx.check(ctx.backend.comm_allreduce_f64(ctx.h, dptr(v), len(v), {"sum": 0, "max": 1, "min": 2}[op]), "comm_allreduce_f64")
    return v


def rebalance_device(popl, tolerance=0.05):
    """ptl_rebalance: returns (n_after, rows sent (+) / received (-))."""
    moved = C.c_int64(0)
    n = int(popl.ctx.backend.rebalance(popl.ctx.h, popl.id, float(tolerance), C.byref(moved)))
    popl.ctx.check(n, "rebalance")
    return n, int(moved.value)


NCOLS = 12      # x0,x1,x2,p0,p1,p2,w,t,s,r (f64) + active (u8) + uid (u64)


def plan_rebalance(counts, tolerance=0.05):
    """Deterministic transfer plan.  counts[r] = particles on rank r.  Returns a list of (src, dst, k): move the
    LAST k rows of src to the end of dst.  Greedy matching of surpluses to deficits in rank order; ranks within
    `tolerance` of the mean are left alone."""
    counts = [int(c) for c in counts]
    n = len(counts)
    total = sum(counts)
    base, extra = divmod(total, n)
    target = [base + (1 if r < extra else 0) for r in range(n)]
    mean = total / n if n else 0
    if mean == 0 or max(abs(c - mean) for c in counts) <= tolerance * mean:
        return []
    surplus = [[r, counts[r] - target[r]] for r in range(n) if counts[r] > target[r]]
    deficit = [[r, target[r] - counts[r]] for r in range(n) if counts[r] < target[r]]
    plan = []
    i = j = 0
    while i < len(surplus) and j < len(deficit):
        k = min(surplus[i][1], deficit[j][1])
        if k > 0:
            plan.append((surplus[i][0], deficit[j][0], k))
        surplus[i][1] -= k
        deficit[j][1] -= k
        if surplus[i][1] == 0:
            i += 1
        if deficit[j][1] == 0:
            j += 1
    return plan


def counts_after(counts, plan):
    out = [int(c) for c in counts]
    for s, d, k in plan:
        out[s] -= k
        out[d] += k
    return out


def gather_counts(dist, n_local, device="cpu"):
    import torch
    world = dist.get_world_size()
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    allc = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allc, mine)
    return [int(c.item()) for c in allc]


def allreduce_diag(dist, vec_sum, vec_max=None):
    """Sum-reduce (and optionally max-reduce) small diagnostic vectors in place."""
    dist.all_reduce(vec_sum, op=dist.ReduceOp.SUM)
    if vec_max is not None:
        dist.all_reduce(vec_max, op=dist.ReduceOp.MAX)
    return vec_sum, vec_max


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def column_views(popl, start, count):
    """torch views of rows [start, start+count) of the 12 device columns of a population (zero copy)."""
    import torch
    views = []
    for col in range(NCOLS):
        typestr, size = ("<f8", 8) if col < 10 else (("|u1", 1) if col == 10 else ("<i8", 8))
        ptr = popl.column_ptr(col) + start * size
        views.append(torch.as_tensor(_DevArray(ptr, count, typestr), device="cuda"))
    return views


def exchange_columns(dist, rank, plan, get_send_views, get_recv_views):
    """Execute a plan with batched point-to-point operations.  `get_send_views(k)` returns the 12 tensors holding the
    last k local rows; `get_recv_views(offset, k)` the 12 tensors where k incoming rows land."""
    ops = []
    recv_offset = 0
    send_taken = 0
    for s, d, k in plan:
        if s == rank:
            for v in get_send_views(send_taken, k):
                ops.append(dist.P2POp(dist.isend, v, d))
            send_taken += k
        elif d == rank:
            for v in get_recv_views(recv_offset, k):
                ops.append(dist.P2POp(dist.irecv, v, s))
            recv_offset += k
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return send_taken, recv_offset


def rebalance(dist, popl, tolerance=0.05):
    """Rebalance one species across the ranks of the default process group.  Returns (n_before, n_after)."""
    import torch
    rank = dist.get_rank()
    n_local = len(popl)
    counts = gather_counts(dist, n_local, device="cuda")
    plan = plan_rebalance(counts, tolerance)
    if not plan:
        return n_local, n_local
    incoming = sum(k for s, d, k in plan if d == rank)
    if n_local + incoming > popl.capacity:
        raise RuntimeError("rebalance: receiving rank lacks capacity")

    def send_views(taken, k):
        return column_views(popl, n_local - taken - k, k)

    def recv_views(offset, k):
        return column_views(popl, n_local + offset, k)

    torch.cuda.synchronize()
    sent, received = exchange_columns(dist, rank, plan, send_views, recv_views)
    torch.cuda.synchronize()
    n_new = n_local - sent + received
    popl.set_n(n_new)
    return n_local, n_new
