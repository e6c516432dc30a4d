"""MultiPopulation and the hot call `advance` — mirrors src/mixed_population.jl.

MultiPopulation :4-18, init! :20-35, advance! :38-47 (advance_init! :97-110 and the advance1! pass loop
:56-93 run inside the library)."""
import ctypes as C

import numpy as np

from ._lib import AdvanceStats, CallbackDesc
from .callback import callback_desc


class MultiPopulation:
    """MultiPopulation(:electron => popl, :photon => popl, ...) — order is the processing order."""

    def __init__(self, *pairs, **named):
        items = list(pairs) + list(named.items())
        self.index = dict(items)
        self.pops = [p for _, p in items]
        self.ctx = self.pops[0].ctx
        ids = (C.c_int32 * len(self.pops))(*[p.id for p in self.pops])
        self.id = self.ctx.check(self.ctx.backend.multipop_create(self.ctx.h, ids, len(self.pops)), "multipop_create")

    def get(self, key):
        """get(mp, ParticleType) :15 — by name or by species id"""
        if key in self.index:
            return self.index[key]
        for p in self.pops:
            if p.species == key:
                return p
        raise KeyError(key)

    def __iter__(self):
        return iter(self.pops)

    def pairs(self):
        return self.index.items()


def init(mpopl):
    """init!(mpopl) :31-33"""
    mpopl.ctx.raise_on_flags(mpopl.ctx.backend.init(mpopl.ctx.h, mpopl.id), "init")


def advance(mpopl, pusher, tfinal, callback=None, check=True):
    """advance!(mpopl, pusher, tfinal, callback) :38-47"""
    ctx = mpopl.ctx
    pd = pusher.desc(ctx)
    cd, walls, counter = callback_desc(callback)
    rc = ctx.backend.advance(ctx.h, mpopl.id, C.byref(pd), float(tfinal), C.byref(cd) if cd is not None else None)
    for iw, w in enumerate(walls):
        n = int(ctx.backend.wall_records(ctx.h, iw, 0, None, None, None, None, 0))
        if n > 0:
            rec = {"x": np.zeros((n, 3)), "p": np.zeros((n, 3)), "w": np.zeros(n), "t": np.zeros(n)}
            ctx.backend.wall_records(ctx.h, iw, n, rec["x"].ctypes.data_as(C.POINTER(C.c_double)),
                                     rec["p"].ctypes.data_as(C.POINTER(C.c_double)),
                                     rec["w"].ctypes.data_as(C.POINTER(C.c_double)),
                                     rec["t"].ctypes.data_as(C.POINTER(C.c_double)), 1)
            w._append(rec)
    if counter is not None:
        counter._harvest(mpopl)
    if check:
        ctx.raise_on_flags(rc, "advance")
    return rc


def last_advance_stats(mpopl):
    st = AdvanceStats()
    mpopl.ctx.backend.last_advance_stats(mpopl.ctx.h, C.byref(st))
    return {k: getattr(st, k) for k, _ in st._fields_}
