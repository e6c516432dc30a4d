"""Collision tables — host-side builders (init code, runs once).

Mirrors src/collision_table.jl (CollisionTable :15-57, ChebyshevCollisionTable :63-106,
collision_table_from_processes :115-149, compratebound :152-167), the `LogLinRange` grid of
src/util.jl:60-127, `ChebContinuumLoss` of src/continuum.jl:25-43 with `energy_loss` :63-139, the
table recipes of scripts/beam.jl:94-129 and the LXCat table layout of src/lxcat.jl:100-134.
The flat arrays produced here are what a Julia host would pass over the C ABI."""
from dataclasses import dataclass, field
import math
import numpy as np

from . import constants as co
from .cheby import BinaryIntervals, chebfit, chebeval, chebdiff
from . import processes as pr
from .processes import ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON, speed


@dataclass
class ChebyshevCollisionTable:
    """collision_table.jl:63-75"""
    proc: list
    b: BinaryIntervals
    rate: np.ndarray        # [order, nprocs, k+1]
    ratebound: np.ndarray   # [order, k+1]
    species: int = ELECTRON
    coefs: list = field(default_factory=list)   # density factor of each (sorted) process

    @property
    def order(self):
        return self.rate.shape[0]

    def __len__(self):
        return len(self.proc)


@dataclass
class CollisionTable:
    """collision_table.jl:15-28 with a LinRange (grid_kind 0) or LogLinRange (grid_kind 1) energy grid
    and a constant rate bound (`maxrate`, collision_table.jl:33) or a vector rate bound on the energy grid
    (`ratebound`, collision_table.jl:35-43)."""
    proc: list
    grid_kind: int
    L1: float
    L2: float
    nE: int
    rate: np.ndarray        # [nprocs, nE]
    maxrate: float
    species: int = SLOW_ELECTRON
    ratebound: np.ndarray = None   # [nE] or None

    def __len__(self):
        return len(self.proc)

    def energy(self):
        L = np.linspace(self.L1, self.L2, self.nE)
        return L if self.grid_kind == 0 else np.exp(L) - math.exp(self.L1)


def loglinrange(L1, L2, N):
    """util.jl:70-77: x = exp(L) - exp(L1) on a linear range L."""
    L = np.linspace(L1, L2, N)
    return np.exp(L) - math.exp(L[0])


def _gamma(species, eng):
    """collision_table.jl:153-156"""
    if species == PHOTON:
        return np.zeros_like(eng)
    return 1 + eng / (co.electron_mass * co.c ** 2)


def compratebound(species, a, b, Fdt, safety):
    """collision_table.jl:152-167: fit of  sum_j f_j + gamma^3 v Fdt |sum_j f_j'| safety."""
    order, nprocs = a.shape[0], a.shape[1]

    def fun(eng):
        if nprocs == 0:
            return np.zeros_like(eng)
        df = sum(chebdiff(eng, b, a[:, j, :]) for j in range(nprocs))
        f = sum(chebeval(eng, b, a[:, j, :]) for j in range(nprocs))
        return f + _gamma(species, eng) ** 3 * speed(species, eng) * Fdt * np.abs(df) * safety

    return chebfit(fun, b, order)


def collision_table_from_processes(processes, species, Fdt, safety=1.1, nintervals=32, order=3,
                                   max_energy=1e3 * co.eV * 2 ** 18):
    """collision_table.jl:115-149.  `processes` is a list of (coef, process)."""
    nprocs = len(processes)
    b = BinaryIntervals(nintervals, max_energy)
    a = np.zeros((order, nprocs, nintervals + 1))
    for j, (coef, proc) in enumerate(processes):
        a[:, j, :] = chebfit(lambda eng, coef=coef, proc=proc: coef * speed(species, eng) * proc.totalcs(eng), b, order)
    rb = compratebound(species, a, b, Fdt, safety)

    eng = loglinrange(math.log(1e-2 * co.eV), math.log(0.9999 * max_energy), 100_000)
    with np.errstate(all="ignore"):
        v1 = speed(species, eng)
        key = np.array([coef * _julia_maximum(proc.totalcs(eng) * v1) for coef, proc in processes])
    # sortperm(..., rev=true): stable, NaN sorts as the largest key
    key = np.where(np.isnan(key), np.inf, key)
    perm = np.argsort(-key, kind="stable")
    proc = [processes[i][1] for i in perm]
    coefs = [processes[i][0] for i in perm]
    return ChebyshevCollisionTable(proc=proc, b=b, rate=np.ascontiguousarray(a[:, perm, :]), ratebound=rb,
                                   species=species, coefs=coefs)


def _julia_maximum(v):
    """Base.maximum propagates NaN."""
    return np.nan if np.any(np.isnan(v)) else np.max(v)


# ---- air tables: scripts/beam.jl:94-129 --------------------------------------------------------
def air_composition(n=co.nair):
    """scripts/beam.jl:39-40"""
    return {"N2": n * 0.79, "O2": n * 0.21}


def build_electron_collision_table(comp, Fdt, safety=1.1, sb=None, **kw):
    """scripts/beam.jl:94-105"""
    from . import seltzer
    sb = sb or {7: seltzer.from_Z(7), 8: seltzer.from_Z(8)}
    processes = [(2 * comp["N2"], pr.RelativisticCoulomb(7)),
                 (2 * comp["O2"], pr.RelativisticCoulomb(8)),
                 (2 * comp["N2"], sb[7]),
                 (2 * comp["O2"], sb[8])]
    processes += [(comp["N2"], orb) for orb in pr.ORBITALS["N2"]]
    processes += [(comp["O2"], orb) for orb in pr.ORBITALS["O2"]]
    return collision_table_from_processes(processes, ELECTRON, Fdt, safety=safety, **kw)


def build_positron_collision_table(comp, tcut, Fdt, safety=1.1, **kw):
    """scripts/beam.jl:107-117"""
    processes = [(2 * comp["N2"], pr.RelativisticCoulomb(7)),
                 (2 * comp["O2"], pr.RelativisticCoulomb(8)),
                 (2 * comp["N2"], pr.Bhaba(7, tcut)),
                 (2 * comp["O2"], pr.Bhaba(8, tcut)),
                 (2 * comp["N2"], pr.PositronAnihilation(7)),
                 (2 * comp["O2"], pr.PositronAnihilation(8))]
    return collision_table_from_processes(processes, POSITRON, Fdt, safety=safety, **kw)


def build_photon_collision_table(comp, *args, **kw):
    """scripts/beam.jl:119-129 (the positional energy argument of the script is ignored there too: Fdt = 0)"""
    processes = [(2 * comp["N2"], pr.PhotoElectric(7)),
                 (2 * comp["O2"], pr.PhotoElectric(8)),
                 (2 * comp["N2"], pr.BetheHeitler(7)),
                 (2 * comp["O2"], pr.BetheHeitler(8)),
                 (2 * comp["N2"], pr.Compton(7)),
                 (2 * comp["O2"], pr.Compton(8))]
    return collision_table_from_processes(processes, PHOTON, 0, **kw)


# ---- continuum losses: src/continuum.jl -----------------------------------------------------------
def _x0x1(C):
    """continuum.jl:123-139"""
    if C < 10:
        return (1.6, 4.0)
    if C < 10.5:
        return (1.7, 4.0)
    if C < 11.0:
        return (1.8, 4.0)
    if C < 11.5:
        return (1.9, 4.0)
    if C < 12.25:
        return (2.0, 4.0)
    if C < 13.804:
        return (2.0, 5.0)
    return (0.326 * C - 2.5, 5.0)


@dataclass
class ContinuumLoss:
    """continuum.jl:6-15"""
    nel: float
    I: float
    Tcut: float

    def energy_loss(self, species, eng):
        """continuum.jl:63-96 (+ taumax :102-103, _F :107-120)"""
        eng = np.asarray(eng, dtype=np.float64)
        nel, I, Tcut = self.nel, self.I, self.Tcut
        mc2, r_e = co.electron_mc2, co.r_e
        with np.errstate(all="ignore"):
            tau = eng / mc2
            tauc = Tcut / mc2
            taumax = tau if species == POSITRON else tau / 2
            g = 1 + tau
            beta2 = 1 - 1 / g ** 2
            tu = np.minimum(tauc, taumax)
            if species == POSITRON:
                y = 1 / (2 + tau)
                F = (np.log(tau * tu) - (tu ** 2 / tau) * (tau * 2 * tu - 3 * tu ** 2 * y / 2 - (tu - tu ** 3 / 3) * y ** 2
                                                           - (tu ** 2 / 2 - tau * tu ** 3 / 3 + tu ** 4 / 4) * y ** 3))
            else:
                F = (-1 - beta2 + np.log((tau - tu) * tu) + tau / (tau - tu)
                     + (tu ** 2 / 2 + (2 * tau + 1) * np.log(1 - tu / tau)) / g ** 2)
            x = np.log(g ** 2 * beta2) / math.log(10) / 2
            hnup = co.hbar * co.c * math.sqrt(4 * math.pi * nel * r_e)
            C = 1 + 2 * math.log(I / hnup)
            xa = C / math.log(10) / 2
            x0, x1 = _x0x1(C)
            a = 2 * math.log(10) * (xa - x) / (x1 - x0) ** 3
            delta = np.where(x < x0, 0.0,
                             np.where(x < x1, 2 * math.log(10) * x - C + a * (x1 - x) ** 3, 2 * math.log(10) * x - C))
            return (2 * math.pi * r_e ** 2 * mc2 * nel / beta2) * (np.log((2 * (g + 1)) / (I / mc2) ** 2) + F - delta)


@dataclass
class ChebContinuumLoss:
    """continuum.jl:25-43"""
    bints: BinaryIntervals
    ec: np.ndarray
    pc: np.ndarray

    @staticmethod
    def from_loss(cl: ContinuumLoss, Tmax, n):
        k = math.ceil(math.log2(Tmax / cl.Tcut))
        bints = BinaryIntervals(k, 2 ** k * cl.Tcut)
        ec = chebfit(lambda x: cl.energy_loss(ELECTRON, x), bints, n)
        pc = chebfit(lambda x: cl.energy_loss(POSITRON, x), bints, n)
        return ChebContinuumLoss(bints, ec, pc)


# ---- LXCat-style linear tables: src/lxcat.jl:100-134 ------------------------------------------------
def lxcat_table_from_rates(procs, nu, grid_kind, L1, L2):
    """Assemble a CollisionTable from per-process collision frequencies nu[nprocs, nE] the way
    load_lxcat does (lxcat.jl:100-134): total rate, maxrate, explicit NullCollision row, processes
    sorted by descending energy-integrated rate."""
    nu = np.asarray(nu, dtype=np.float64)
    nprocs, nE = nu.shape
    rate = np.zeros((nprocs + 1, nE))
    rate[:nprocs] = nu
    nutotal = rate.sum(axis=0)
    maxrate = float(nutotal.max())
    rate[nprocs] = maxrate - nutotal
    procs = list(procs) + [pr.NullCollision()]
    inteng = rate.sum(axis=1)
    perm = np.argsort(-inteng, kind="stable")
    return CollisionTable(proc=[procs[i] for i in perm], grid_kind=grid_kind, L1=L1, L2=L2, nE=nE,
                          rate=np.ascontiguousarray(rate[perm]), maxrate=maxrate)


def synthetic_lxcat_table(n=co.nair, nE=4096, emax_eV=100.0, grid_kind=0, extra_levels=0):
    """synthetic cross-section set (the reference ships no LXCat data): elastic
    (1e-19 m^2, mass ratio 1.95e-5), three excitations (0.3 / 6.2 / 11 eV), ionisation (15.6 eV) and a
    dissociative-attachment resonance peaked at 6.5 eV.  `extra_levels` adds that many further excitation
    channels (thresholds 0.2, 0.45, ... eV): a real N2/O2 set has 50-80 channels."""
    if grid_kind == 0:
        L1, L2 = 0.0, emax_eV * co.eV
        eng = np.linspace(L1, L2, nE)
    else:
        L1, L2 = math.log(1e-3 * co.eV), math.log(emax_eV * co.eV)
        eng = loglinrange(L1, L2, nE)
    e = eng / co.eV
    v = np.sqrt(2 * eng / co.electron_mass)

    def thr(e0, s0, width):
        x = np.maximum(e - e0, 0.0)
        return s0 * x / (x + width) * (1.0 / (1.0 + x / 50.0))

    sig = [np.full_like(e, 1e-19),
           thr(0.3, 2.0e-21, 1.0), thr(6.2, 6.0e-21, 4.0), thr(11.0, 1.2e-20, 6.0),
           thr(15.6, 2.5e-20, 30.0),
           1.5e-22 * np.exp(-0.5 * ((e - 6.5) / 1.0) ** 2)]
    procs = [pr.Elastic(1.95e-5), pr.Excitation(0.3 * co.eV), pr.Excitation(6.2 * co.eV), pr.Excitation(11.0 * co.eV),
             pr.Ionization(15.6 * co.eV), pr.Attachment(0.0)]
    for k in range(extra_levels):
        t0 = 0.2 + 0.25 * k
        sig.append(thr(t0, 2.0e-22, 3.0))
        procs.append(pr.Excitation(t0 * co.eV))
    nu = np.stack([s * v * n for s in sig])
    return lxcat_table_from_rates(procs, nu, grid_kind, L1, L2)
