"""Callbacks — mirrors src/callback.jl.

In-loop hooks (`onadvance`, `oncollision`) cannot call back into the host from a GPU kernel, so only
the stock callbacks that the reference ships are supported inside `advance`: VoidCallback :9,
WallCallback :146-184, CollisionCounter :118-141, and CombinedCallback :41-108 of those.  The
between-step hooks (`onstep`, `onoutput`) run on the host exactly as in the reference:
ParticleCountCallback :190-204, RouletteCallback :210-226, SplitCallback :231-247,
PopulationTargetCallback :253-268."""
import numpy as np

from ._lib import CallbackDesc, MAX_WALLS
from . import population as _pop


class AbstractCallback:
    def onstep(self, mpopl, t=None):
        return True

    def onoutput(self, mpopl, t=None, i=None):
        return None

    def _collect(self, walls, flags):
        pass


class VoidCallback(AbstractCallback):
    pass


class CombinedCallback(AbstractCallback):
    def __init__(self, tpl):
        self.tpl = tuple(tpl)

    def __getitem__(self, i):
        return self.tpl[i]

    def onstep(self, mpopl, t=None):
        r = True
        for c in self.tpl:           # callback.jl:101-106: all are called, results and-ed
            r = c.onstep(mpopl, t) and r
        return r

    def onoutput(self, mpopl, t=None, i=None):
        for c in self.tpl:
            c.onoutput(mpopl, t, i)

    def _collect(self, walls, flags):
        for c in self.tpl:
            c._collect(walls, flags)


class CollisionCounter(AbstractCallback):
    """callback.jl:118-141.  Counts are kept per process on the device and aggregated per process
    type here, like the reference's Dict{Type,Int}."""

    def __init__(self):
        self.d = {}

    def _collect(self, walls, flags):
        flags["count"] = self

    def _harvest(self, mpopl):
        for popl in mpopl:
            counts = popl.collision_counts(clear=True)
            names = [p.name for p in popl.collisions.proc] + ["NullCollision"]
            for nm, c in zip(names, counts):
                if c:
                    self.d[nm] = self.d.get(nm, 0) + int(c)

    def __repr__(self):
        lines = ["CollisionCounter with values:"] + [f"{k}:   {v}" for k, v in self.d.items()]
        return "\n".join(lines + [f"Total:   {sum(self.d.values())}"])


class WallCallback(AbstractCallback):
    """callback.jl:146-184: records the interpolated state of particles of `species` crossing the plane
    x[coord] = v in the + direction; `drop` deactivates them.  coord is 1-based like the reference."""

    def __init__(self, species, coord, v, drop=True):
        self.species, self.coord, self.v, self.drop = species, coord, v, drop
        self.accum = {"x": np.zeros((0, 3)), "p": np.zeros((0, 3)), "w": np.zeros(0), "t": np.zeros(0)}

    def _collect(self, walls, flags):
        walls.append(self)

    def _append(self, rec):
        for k in self.accum:
            self.accum[k] = np.concatenate([self.accum[k], rec[k]])


def callback_desc(cb):
    """Flatten a callback tree into the C descriptor; returns (desc or None, walls, counter)."""
    if cb is None or type(cb) is VoidCallback:
        return None, [], None
    walls, flags = [], {}
    cb._collect(walls, flags)
    if not walls and "count" not in flags:
        return None, [], None
    if len(walls) > MAX_WALLS:
        raise ValueError(f"at most {MAX_WALLS} WallCallbacks")
    d = CallbackDesc()
    d.nwalls = len(walls)
    d.count_collisions = 1 if "count" in flags else 0
    for i, w in enumerate(walls):
        d.wall[i].species = w.species
        d.wall[i].coord = w.coord - 1
        d.wall[i].v = float(w.v)
        d.wall[i].drop = 1 if w.drop else 0
    return d, walls, flags.get("count")


class ParticleCountCallback(AbstractCallback):
    """callback.jl:190-204"""

    def __init__(self, particles):
        self.p = list(particles)
        self.counts = []

    def onoutput(self, mpopl, t=None, i=None):
        self.counts.append([t] + [_pop.weight(mpopl.get(s)) for s in self.p])


class RouletteCallback(AbstractCallback):
    """callback.jl:210-226"""

    def __init__(self, m):
        self.m = m

    def onstep(self, mpopl, t=None):
        for popl in mpopl:
            n = _pop.nactives(popl)
            if n > self.m:
                _pop.roulette(self.m / n, popl)
            _pop.repack(popl)
        return True


class SplitCallback(AbstractCallback):
    """callback.jl:231-247"""

    def __init__(self, species, m, min_frac):
        self.species, self.m, self.min_frac = species, m, min_frac

    def onstep(self, mpopl, t=None):
        popl = mpopl.get(self.species)
        n = _pop.nactives(popl)
        if 0 < n < self.m * self.min_frac:
            _pop.split(self.m / n - 1, popl)
        _pop.repack(popl)
        return True


class PopulationTargetCallback(AbstractCallback):
    """callback.jl:253-268"""

    def __init__(self, species, n, ismax=True):
        self.species, self.n, self.ismax = species, n, ismax

    def onstep(self, mpopl, t=None):
        na = _pop.nactives(mpopl.get(self.species))
        return na < self.n if self.ismax else na > self.n
