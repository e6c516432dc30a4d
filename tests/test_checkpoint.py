"""Checkpoint / restart: k steps + save + restore into a fresh context + k steps must equal 2k
uninterrupted steps, row for row and bit for bit (the RNG is counter-based: uid, seed and the advance-call index are all
that has to survive).  CPU leg through the oracle backend, GPU leg through the CUDA library."""
import io

import numpy as np
import pytest

import particulator_b200 as P
from particulator_b200 import checkpoint
from conftest import make_world, default_pusher
from oracle_backend import oracle_context

DT = 2.5e-11


def _steps(mp, pops, t, k):
    for _ in range(k):
        t += DT
        P.advance(mp, default_pusher(), t)
        for q in pops:
            P.droplow(q)
    return t


def _roundtrip(make_ctx, air_tables, tmp_path):
    a = make_ctx()
    a.set_rng(7, 0)
    mpa, *pa = make_world(a, air_tables, 600, 400, 50, cap=20000, seed=4)
    ta = _steps(mpa, pa, 0.0, 2)
    path = str(tmp_path / "state.ckpt")         # no .npz suffix: the name must be used verbatim (np.savez would append one)
    meta = checkpoint.save_checkpoint(path, mpa, ta, extra={"note": "after 2 steps"})
    assert (tmp_path / "state.ckpt").exists()
    assert meta["step"] == a.get_rng()[1] and meta["seed"] == 7
    assert meta["next_uid"] == a.get_uid_counter() > 1
    blob = checkpoint.dumps(mpa, ta)
    ta = _steps(mpa, pa, ta, 2)                 # uninterrupted run continues

    b = make_ctx()                              # a fresh context, populations rebuilt empty by "the script"
    b.set_rng(12345, 99)                        # wrong on purpose: the checkpoint must overwrite it
    mpb, *pb = make_world(b, air_tables, 0, 0, 0, cap=20000, seed=0)
    tb = checkpoint.load_checkpoint(path, mpb)
    assert tb == pytest.approx(2 * DT) and b.get_rng() == (7, meta["step"])
    # the uid counter survives: a particle injected after the restore gets a uid no live particle has (uids key the RNG)
    assert b.get_uid_counter() >= meta["next_uid"]
    live = np.concatenate([q.download()["uid"] for q in pb])
    seq = live[(live >> np.uint64(63)) == 0]
    assert b.get_uid_counter() > int(seq.max())
    tb = _steps(mpb, pb, tb, 2)
    assert ta == tb
    for qa, qb in zip(pa, pb):
        da, db = qa.download(), qb.download()
        oa, ob = np.argsort(da["uid"], kind="stable"), np.argsort(db["uid"], kind="stable")
        assert len(da["uid"]) == len(db["uid"])
        for c in da:
            np.testing.assert_array_equal(da[c][oa], db[c][ob], err_msg=c)
    # bytes form restores the same state as the file form
    c = make_ctx()
    mpc, *pc = make_world(c, air_tables, 0, 0, 0, cap=20000, seed=0)
    checkpoint.loads(blob, mpc)
    m2, st = checkpoint.read_checkpoint(io.BytesIO(blob))
    for q, pm in zip(pc, m2["populations"]):
        assert len(q) == pm["n"]
        np.testing.assert_array_equal(q.download()["p"], st[pm["name"]]["p"])
    for ctx in (a, b, c):
        ctx.close()


def test_checkpoint_restart_oracle(air_tables, tmp_path):
    _roundtrip(oracle_context, air_tables, tmp_path)


@pytest.mark.gpu
def test_checkpoint_restart_gpu(air_tables, tmp_path):
    _roundtrip(lambda: P.Context(device=0), air_tables, tmp_path)


def test_checkpoint_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.npz"
    np.savez(p, meta=np.frombuffer(b'{"format": "other"}', dtype=np.uint8))
    with pytest.raises(ValueError):
        checkpoint.read_checkpoint(str(p))


def test_restore_without_saved_counter_still_avoids_live_uids(air_tables):
    """Explicit uids move the default-uid counter past the largest sequential uid (a restored context used to
    restart at 1 + sum(n) and could reissue a live uid)."""
    ctx = oracle_context()
    mp, el, ph, po = make_world(ctx, air_tables, 50, 0, 0, cap=1000, seed=1)    # uids 1..50 (explicit, conftest)
    assert ctx.get_uid_counter() == 51
    j = P.add_particle(el, [0, 0, 0], [0, 0, 3e-22], uid=0)
    assert j == 50 and int(el.download()["uid"][50]) == 51
    ctx.close()


def test_load_rejects_populations_missing_from_the_file(air_tables, tmp_path):
    a = oracle_context()
    mpa, ela, pha, poa = make_world(a, air_tables, 10, 0, 0, cap=100, seed=1)
    one = P.MultiPopulation(("electron", ela))
    path = str(tmp_path / "only_e.ckpt")
    checkpoint.save_checkpoint(path, one, 0.0)
    with pytest.raises(KeyError):
        checkpoint.load_checkpoint(path, mpa)      # photon / positron populations are not in the file
    a.close()
