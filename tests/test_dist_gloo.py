"""Host-side logic of the multi-GPU path on CPU: world_size-2 `gloo` process group (no GPU needed)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_plan_rebalance_properties():
    from particulator_b200.dist import plan_rebalance, counts_after
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        counts = [int(c) for c in rng.integers(0, 10000, n)]
        plan = plan_rebalance(counts, tolerance=0.0)
        after = counts_after(counts, plan)
        assert sum(after) == sum(counts)
        assert max(after) - min(after) <= 1
        assert all(k > 0 and s != d for s, d, k in plan)
        assert all(a >= 0 for a in after)
        # a rank never both sends and receives
        assert not ({s for s, _, _ in plan} & {d for _, d, _ in plan})
    assert plan_rebalance([100, 101, 99, 100], tolerance=0.05) == []
    assert plan_rebalance([0, 0, 0], tolerance=0.05) == []
    assert plan_rebalance([10, 0]) == [(0, 1, 5)]


def test_c_abi_plan_equals_python_plan():
    """ptl_rebalance_plan (the plan ptl_rebalance executes over NCCL, pure host code in the CUDA library) against the Python
    statement of the same greedy matching."""
    import ctypes as C
    import particulator_b200 as P
    from particulator_b200.dist import plan_rebalance
    dll = C.CDLL(P.LIB_PATH)
    f = dll.ptl_rebalance_plan
    f.restype = C.c_int32
    f.argtypes = [C.POINTER(C.c_int64), C.c_int32, C.c_double, C.POINTER(C.c_int64), C.c_int32]
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(1, 17))
        counts = np.ascontiguousarray(rng.integers(0, 10 ** int(rng.integers(1, 9)), n), dtype=np.int64)
        tol = float(rng.choice([0.0, 0.05, 0.5]))
        moves = np.zeros(3 * n, dtype=np.int64)
        m = f(counts.ctypes.data_as(C.POINTER(C.c_int64)), n, tol, moves.ctypes.data_as(C.POINTER(C.c_int64)), n)
        assert m >= 0
        got = [tuple(int(v) for v in moves[3 * q:3 * q + 3]) for q in range(m)]
        assert got == plan_rebalance([int(c) for c in counts], tol), (counts, tol)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from particulator_b200.dist import gather_counts, plan_rebalance, exchange_columns, allreduce_diag, NCOLS
        # a fake shard: 12 columns of a "population" held in CPU tensors
        n_local = 1000 if rank == 0 else 200
        cap = 2000
        cols = [torch.zeros(cap, dtype=torch.float64) for _ in range(10)] + [torch.zeros(cap, dtype=torch.uint8), torch.zeros(cap, dtype=torch.int64)]
        for c in range(10):
            cols[c][:n_local] = torch.arange(n_local, dtype=torch.float64) + 10000 * rank + 0.001 * c
        cols[10][:n_local] = 1
        cols[11][:n_local] = torch.arange(n_local) + 10 ** 6 * (rank + 1)
        counts = gather_counts(dist, n_local)
        assert counts == [1000, 200]
        plan = plan_rebalance(counts)
        assert plan == [(0, 1, 400)]
        sent, recvd = exchange_columns(dist, rank, plan,
                                       lambda taken, k: [c[n_local - taken - k:n_local - taken] for c in cols],
                                       lambda off, k: [c[n_local + off:n_local + off + k] for c in cols])
        n_new = n_local - sent + recvd
        assert n_new == 600
        # every uid still exists exactly once across the two ranks, and rows stayed intact (column c = base + 0.001 c)
        uids = cols[11][:n_new].clone()
        alluids = [torch.zeros(600, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(alluids, uids)
        cat = torch.cat(alluids)
        assert len(torch.unique(cat)) == 1200
        base = cols[0][:n_new]
        for c in range(10):
            assert torch.allclose(cols[c][:n_new], base + 0.001 * c, rtol=0, atol=1e-9)
        assert int(cols[10][:n_new].sum()) == n_new
        # diagnostics reduction: global count / weight / max energy
        vs = torch.tensor([float(n_new), 2.0 * n_new], dtype=torch.float64)
        vm = torch.tensor([float(rank + 1)], dtype=torch.float64)
        allreduce_diag(dist, vs, vm)
        assert vs.tolist() == [1200.0, 2400.0] and vm.item() == 2.0
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_rebalance_and_diag_reduce_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
