"""Population of one particle species — host handle of a device-resident SoA store.

Mirrors src/population.jl: Population :7-44, nparticles :78, nactives :89, add_particle! :103,
remove_particle! :120, weight :130, meanenergy :152, maxenergy :172, spread :180, posvar :204,
repack! :229, droplow! :273, roulette! :291, split! :316.  Julia's `f!` is spelled `f` here."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import PtlError, DiagOut, dptr, as_f64
from . import constants as co
from .processes import ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON

_COLS = ("x", "p", "w", "t", "s", "r", "active", "uid")


def kinenergy(species, p):
    """electron.jl:53, positron.jl:40, photon.jl:52, slow-electron.jl:27 (host-side convenience)."""
    p = np.asarray(p, dtype=np.float64)
    p2 = np.sum(p * p, axis=-1)
    if species == PHOTON:
        return np.sqrt(p2) * co.c
    if species == SLOW_ELECTRON:
        return 0.5 * co.electron_mass * p2
    return np.sqrt(co.electron_mc2 ** 2 + co.c ** 2 * p2) - co.electron_mc2


def momentum_norm_from_kin(species, kin):
    """electron.jl:51, positron.jl:38, photon.jl:48"""
    if species == PHOTON:
        return kin / co.c
    if species == SLOW_ELECTRON:
        return np.sqrt(2 * kin / co.electron_mass)
    return np.sqrt((kin + co.electron_mc2) ** 2 - co.electron_mc2 ** 2) / co.c


class Population:
    """Population(max_particles, inparticles, collisions, energy_cut)  population.jl:34-44.

    `init` is a dict of arrays: x[n,3], p[n,3] and optionally w, t, s, r, active, uid.  Missing `s`
    is drawn as -log(u) (the reference's constructor default `s = nextcoll()`, electron.jl:30)."""

    def __init__(self, ctx, species, max_particles, init, collisions, energy_cut=0.0, rng=None):
        self.ctx, self.species, self.collisions, self.energy_cut = ctx, species, collisions, float(energy_cut)
        self.table_id = ctx.table(collisions)
        self.capacity = int(max_particles)
        self.id = ctx.check(ctx.backend.population_create(ctx.h, species, self.capacity, self.energy_cut, self.table_id),
                            "population_create")
        if init is not None and len(init.get("x", ())) > 0:
            self.upload(init, rng=rng)

    # -- bulk transfer ----------------------------------------------------------------------------
    def upload(self, st, rng=None):
        x = as_f64(st["x"]).reshape(-1, 3)
        n = x.shape[0]
        p = as_f64(st["p"]).reshape(n, 3)
        w = as_f64(st.get("w", np.ones(n)))
        t = as_f64(st.get("t", np.zeros(n)))
        if "s" in st:
            s = as_f64(st["s"])
        else:
            rng = rng or np.random.default_rng(0)
            s = -np.log(1.0 - rng.random(n))
        r = as_f64(st.get("r", np.zeros(n)))
        active = np.ascontiguousarray(st.get("active", np.ones(n, dtype=np.uint8)), dtype=np.uint8)
        uid = st.get("uid")
        if uid is not None:
            uid = np.ascontiguousarray(uid, dtype=np.uint64)
        b = self.ctx.backend
        rc = b.population_upload(self.ctx.h, self.id, n, dptr(x), dptr(p), dptr(w), dptr(t), dptr(s), dptr(r),
                                 active.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 None if uid is None else uid.ctypes.data_as(C.POINTER(C.c_uint64)))
        self.ctx.check(rc, "population_upload")

    def download(self, columns=_COLS):
        n = len(self)
        out = {}
        if "x" in columns: out["x"] = np.zeros((n, 3))
        if "p" in columns: out["p"] = np.zeros((n, 3))
        for k in ("w", "t", "s", "r"):
            if k in columns: out[k] = np.zeros(n)
        if "active" in columns: out["active"] = np.zeros(n, dtype=np.uint8)
        if "uid" in columns: out["uid"] = np.zeros(n, dtype=np.uint64)
        b = self.ctx.backend
        got = b.population_download(self.ctx.h, self.id, n, dptr(out.get("x")), dptr(out.get("p")), dptr(out.get("w")),
                                    dptr(out.get("t")), dptr(out.get("s")), dptr(out.get("r")),
                                    None if "active" not in out else out["active"].ctypes.data_as(C.POINTER(C.c_uint8)),
                                    None if "uid" not in out else out["uid"].ctypes.data_as(C.POINTER(C.c_uint64)))
        self.ctx.check(int(got), "population_download")
        return out

    def __len__(self):
        return int(self.ctx.check(int(self.ctx.backend.population_n(self.ctx.h, self.id)), "population_n"))

    def column_ptr(self, col):
        return self.ctx.backend.population_column_ptr(self.ctx.h, self.id, col)

    def set_n(self, n):
        self.ctx.check(self.ctx.backend.population_set_n(self.ctx.h, self.id, int(n)), "population_set_n")

    def diag(self):
        d = DiagOut()
        self.ctx.check(self.ctx.backend.diag(self.ctx.h, self.id, C.byref(d)), "diag")
        return d

    def collision_counts(self, clear=False):
        cnt = np.zeros(len(self.collisions.proc) + 1, dtype=np.int64)
        self.ctx.check(self.ctx.backend.collision_counts(self.ctx.h, self.table_id, cnt.ctypes.data_as(C.POINTER(C.c_int64)),
                                                         1 if clear else 0), "collision_counts")
        return cnt

    def histogram(self, quantity, lo, hi, nbins, logscale=False):
        out = np.zeros(nbins)
        q = {"energy": 0, "costheta": 1}[quantity]
        self.ctx.check(self.ctx.backend.histogram(self.ctx.h, self.id, q, float(lo), float(hi), nbins, 1 if logscale else 0,
                                                  dptr(out)), "histogram")
        return out


# ---- generic functions of the reference (population.jl) ---------------------------------------------
def nparticles(popl):
    return len(popl)


def nactives(popl):
    return int(popl.diag().nactive)


def weight(popl):
    return popl.diag().weight


def meanenergy(popl):
    d = popl.diag()
    return d.wenergy / d.weight if d.weight != 0 else math.nan


def maxenergy(popl):
    return popl.diag().maxenergy


def spread(popl):
    d = popl.diag()
    if d.weight == 0:
        return np.full(3, math.nan), math.nan
    xm = np.array(d.wx[:]) / d.weight
    x2 = d.wr2 / d.weight
    return xm, math.sqrt(abs(x2 - float(xm @ xm)))


def posvar(popl):
    d = popl.diag()
    return np.array(d.wx2[:]) / d.weight - (np.array(d.wx[:]) / d.weight) ** 2


def empty(popl):
    popl.ctx.check(popl.ctx.backend.population_clear(popl.ctx.h, popl.id), "population_clear")


def add_particle(popl, x, p, w=1.0, t=0.0, s=None, r=0.0, uid=0):
    s = -math.log(1.0 - np.random.random()) if s is None else s
    x, p = as_f64(x), as_f64(p)
    j = int(popl.ctx.backend.population_append(popl.ctx.h, popl.id, dptr(x), dptr(p), w, t, s, r, int(uid)))
    if j == -7:       # PTL_ECAPACITY: the reference asserts n < length(particles) (population.jl:107)
        raise PtlError("add_particle!: population is full (CAPACITY_OVERFLOW)")
    if j < -1:        # -1 alone means "below the energy cut, not added" (population.jl:105)
        popl.ctx.check(j, "population_append")
    return j


def remove_particle(popl, i):
    popl.ctx.check(popl.ctx.backend.population_deactivate(popl.ctx.h, popl.id, int(i)), "population_deactivate")


def repack(popl):
    return int(popl.ctx.check(int(popl.ctx.backend.repack(popl.ctx.h, popl.id)), "repack"))


def droplow(popl, thres=0.0):
    return int(popl.ctx.check(int(popl.ctx.backend.droplow(popl.ctx.h, popl.id, float(thres))), "droplow"))


def _law_nodes(f, lo, hi, nodes, logscale):
    """Sample a Python callable f(energy [J]) on the nodes the library interpolates between."""
    x = np.linspace(math.log10(lo) if logscale else lo, math.log10(hi) if logscale else hi, nodes)
    e = 10.0 ** x if logscale else x
    v = np.ascontiguousarray([float(f(float(q))) for q in e], dtype=np.float64)
    return float(x[0]), float(x[-1]), v


def roulette(p, popl, lo=None, hi=None, nodes=1024, logscale=True):
    """roulette!(f_or_p, popl) (population.jl:291-314).  A number is the constant retain probability; a callable f(energy)
    is sampled on `nodes` nodes between `lo` and `hi` (default: energy_cut .. 1 GeV, log-spaced) and interpolated linearly
    on the device — a closure cannot cross the C ABI."""
    if callable(p):
        lo = lo if lo is not None else max(popl.energy_cut, 1e-3 * co.eV)
        hi = hi if hi is not None else 1e9 * co.eV
        x0, x1, v = _law_nodes(p, lo, hi, nodes, logscale)
        rc = popl.ctx.backend.roulette_law(popl.ctx.h, popl.id, x0, x1, len(v), 1 if logscale else 0, dptr(v))
    else:
        rc = popl.ctx.backend.roulette(popl.ctx.h, popl.id, float(p))
    popl.ctx.raise_on_flags(rc, "roulette")


def split(p, popl, lo=None, hi=None, nodes=1024, logscale=True):
    """split!(f_or_p, popl) (population.jl:316-340): mean number of copies, constant or a callable of the energy."""
    if callable(p):
        lo = lo if lo is not None else max(popl.energy_cut, 1e-3 * co.eV)
        hi = hi if hi is not None else 1e9 * co.eV
        x0, x1, v = _law_nodes(p, lo, hi, nodes, logscale)
        rc = popl.ctx.backend.split_law(popl.ctx.h, popl.id, x0, x1, len(v), 1 if logscale else 0, dptr(v))
    else:
        rc = popl.ctx.backend.split(popl.ctx.h, popl.id, float(p))
    popl.ctx.raise_on_flags(rc, "split")


def shuffle(popl):
    """shuffle!(popl) (population.jl:266-271): a uniformly distributed permutation of the rows."""
    popl.ctx.check(popl.ctx.backend.shuffle(popl.ctx.h, popl.id), "shuffle")
