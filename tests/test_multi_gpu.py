"""Multi-GPU entry points of the C ABI on real devices (skipped with fewer than two GPUs; run with `gpurun --gpus 2`).

  * contexts on two devices in ONE process (kernel attributes are per device — the second device used to fail),
  * ptl_comm_init / ptl_diag_allreduce / ptl_histogram_allreduce / ptl_rebalance over NCCL, two ranks as two host threads of
    one process (the library binds the calling thread to its context's device), uid uniqueness and weight / energy
    conservation across the exchange."""
import threading

import numpy as np
import pytest

import particulator_b200 as P
from particulator_b200 import dist as pdist
from conftest import make_world, default_pusher

pytestmark = pytest.mark.gpu
co = P.co


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")


@needs2
def test_contexts_on_two_devices_in_one_process(air_tables):
    out = []
    for dev in (0, 1):
        ctx = P.Context(device=dev)
        ctx.set_option("small_pass_rows", 0)          # the warp-private kernel needs its shared-memory opt-in on THIS device
        ctx.set_rng(3, 0)
        mp, el, ph, po = make_world(ctx, air_tables, 3000, 1000, 300, cap=40000, seed=8)
        P.advance(mp, default_pusher(), 2.5e-11)
        d = el.download()
        o = np.argsort(d["uid"], kind="stable")
        out.append({k: v[o] for k, v in d.items()})
        assert ctx.error_flags() == 0
        ctx.close()
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k]), k      # same inputs, same seed: the device does not matter


@needs2
def test_comm_allreduce_and_rebalance_two_ranks(air_tables):
    nranks = 2
    ids = {}
    barrier = threading.Barrier(nranks)
    res = [None] * nranks
    errs = []

    def exchange(rank):
        def f(payload):
            if rank == 0:
                ids["id"] = payload
            barrier.wait()
            return ids["id"]
        return f

    def work(rank):
        try:
            ctx = P.Context(device=rank)
            pdist.init_comm(ctx, rank=rank, nranks=nranks, exchange=exchange(rank))
            ctx.set_rng(11, 0)
            n = 20000 if rank == 0 else 4000
            mp, el, ph, po = make_world(ctx, air_tables, n, 0, 0, cap=60000, seed=100 + rank)
            # disjoint uids: rank r starts at r * 2^40 + 1 (ptl_comm_init moved the counter); re-upload without explicit uids
            d = el.download(); d.pop("uid"); el.upload(d)
            before = el.download()
            loc = el.diag()
            glob = pdist.diag_allreduce(el)
            h_loc = el.histogram("energy", 1e3 * co.eV, 1e8 * co.eV, 32, logscale=True)
            h_glob = pdist.histogram_allreduce(el, "energy", 1e3 * co.eV, 1e8 * co.eV, 32, logscale=True)
            n_after, moved = pdist.rebalance_device(el, tolerance=0.05)
            after = el.download()
            glob2 = pdist.diag_allreduce(el)
            res[rank] = dict(loc=(loc.n, loc.weight, loc.wenergy), glob=(glob.n, glob.weight, glob.wenergy, glob.maxenergy),
                             glob2=(glob2.n, glob2.weight, glob2.wenergy), h_loc=h_loc, h_glob=h_glob, n_after=n_after, moved=moved,
                             before=before, after=after)
            barrier.wait()
            pdist.destroy_comm(ctx)
            ctx.close()
        except Exception as exc:  # pragma: no cover
            errs.append(repr(exc))
            try:
                barrier.abort()
            except Exception:
                pass

    ths = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=300)
    assert not errs, errs
    a, b = res
    # global diagnostics = sum of the local ones, identical on both ranks
    assert a["glob"] == b["glob"]
    assert a["glob"][0] == a["loc"][0] + b["loc"][0] == 24000
    assert a["glob"][1] == pytest.approx(a["loc"][1] + b["loc"][1], rel=1e-13)
    assert a["glob"][2] == pytest.approx(a["loc"][2] + b["loc"][2], rel=1e-13)
    np.testing.assert_allclose(a["h_glob"], a["h_loc"] + b["h_loc"], rtol=1e-13)
    assert np.array_equal(a["h_glob"], b["h_glob"])
    # rebalance: 20000 / 4000 -> 12000 / 12000, 8000 rows moved from rank 0 to rank 1
    assert (a["n_after"], b["n_after"]) == (12000, 12000) and a["moved"] == 8000 and b["moved"] == -8000
    uids = np.concatenate([a["after"]["uid"], b["after"]["uid"]])
    uids0 = np.concatenate([a["before"]["uid"], b["before"]["uid"]])
    assert len(np.unique(uids)) == len(uids) == 24000 and np.array_equal(np.sort(uids), np.sort(uids0))
    # every column travelled with its row
    src = {int(u): i for i, u in enumerate(a["before"]["uid"])}
    moved_rows = b["after"]["uid"][4000:]
    idx = np.array([src[int(u)] for u in moved_rows])
    for k in ("x", "p", "w", "t", "s", "r", "active"):
        assert np.array_equal(b["after"][k][4000:], a["before"][k][idx]), k
    assert a["glob2"][0] == 24000 and a["glob2"][1] == pytest.approx(a["glob"][1], rel=1e-13) and a["glob2"][2] == pytest.approx(a["glob"][2], rel=1e-12)
