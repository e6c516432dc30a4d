"""LXCat cross-section database reader: the host-init builder that feeds the linear-table path of the kernels
(reference `load_lxcat`, src/lxcat.jl:17-134, and `ensure_elastic`, :147-158).

Input: one or more JSON files, each a list of records
    {"target": "N2", "kind": "ELASTIC" | "EFFECTIVE" | "EXCITATION" | "IONIZATION" | "ATTACHMENT",
     "data": [[energy_eV, sigma_m2], ...], "comment": "...", "threshold": eV, "mass_ratio": m/M,
     optional "rescale", "weight_scale", "product"}
Output: collision frequencies nu = density * v(E) * sigma(E) on the caller's energy grid, one row per process plus the
explicit NullCollision row `maxrate - sum`, rows sorted by descending energy-summed rate — the `rate[nprocs+1, nE]`
array `ptl_table_create_linear` takes.

Differences from the reference, on purpose:
  * the reference walks `targets::Dict()` in Julia's hash order before the final sort; here targets are walked in
    first-appearance order.  The final order is decided by the sort (stable, ties keep this pre-order), so the two can
    only differ for processes with exactly equal summed rates;
  * "PHOTOEMISSION" records raise: PhotoEmission is outside the scoped path.
"""
import json
import math

import numpy as np

from . import constants as co
from . import processes as pr
from .tables import CollisionTable

_KINDS = {
    "ATTACHMENT": lambda itm: pr.Attachment(itm["threshold"] * co.eV),      # lxcat.jl:70-71
    "EXCITATION": lambda itm: pr.Excitation(itm["threshold"] * co.eV),      # :73-74
    "IONIZATION": lambda itm: pr.Ionization(itm["threshold"] * co.eV),      # :76-77
    "ELASTIC": lambda itm: pr.Elastic(itm["mass_ratio"]),                   # :79-80
}


def _flat_linear(x0, y0, x):
    """Gridded-linear interpolation with flat extrapolation (Interpolations.jl `extrapolate(interpolate((x0,), y0,
    Gridded(Linear())), Flat())`, lxcat.jl:63-64).  Knots must ascend strictly, as Interpolations requires."""
    x0 = np.asarray(x0, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    if len(x0) < 2 or np.any(np.diff(x0) <= 0):
        raise ValueError("cross-section knots must be at least two strictly ascending energies")
    xc = np.clip(x, x0[0], x0[-1])
    k = np.clip(np.searchsorted(x0, xc, side="right") - 1, 0, len(x0) - 2)
    f = (xc - x0[k]) / (x0[k + 1] - x0[k])
    return (1 - f) * y0[k] + f * y0[k + 1]


def ensure_elastic(procs):
    """A target given with an EFFECTIVE (momentum-transfer) cross-section gets ELASTIC = EFFECTIVE - sum of the
    target's excitations and ionisations (lxcat.jl:147-158)."""
    for p in procs:
        if p["kind"] != "EFFECTIVE":
            continue
        for p2 in procs:
            if p2 is not p and p2["kind"] in ("EXCITATION", "IONIZATION"):
                p["nu"] = p["nu"] - p2["nu"]
        p["kind"] = "ELASTIC"


def signature(item):
    return f"e + {item['target']} -> {item.get('product', '')}... ({item['kind']})"


def load_lxcat(fnames, densities, energy, photon_weight=1.0, photon_multiplier=1.0, verbose=False):
    """Returns a dict with the reference's named-tuple fields `proc`, `rate`, `maxrate`, `origperm`
    (lxcat.jl:133).  `energy` is the grid in joules; `densities` maps target name -> number density (m^-3)."""
    if isinstance(fnames, (str, bytes)):
        fnames = [fnames]
    db = []
    for fname in fnames:                                        # :23-30
        with open(fname, "r") as fd:
            db.extend(json.load(fd))
    energy = np.asarray(energy, dtype=np.float64)
    v = np.sqrt(2 * energy / co.electron_mass)                  # :36
    targets = {}
    nprocs = 0
    for item in db:
        dens = float(densities.get(item["target"], 0.0))        # :40-41
        if dens == 0.0:
            continue
        if item["kind"] == "PHOTOEMISSION":
            raise NotImplementedError("PHOTOEMISSION records: PhotoEmission is outside the scoped hot path")
        item = dict(item)
        energy0 = np.array([d[0] * co.eV for d in item["data"]], dtype=np.float64)
        cs0 = np.array([d[1] for d in item["data"]], dtype=np.float64)
        if "3-body" in item.get("comment", ""):                 # :47-49, three-body attachment: sigma ∝ density (cm^-3)
            cs0 = cs0 * (dens / co.centi ** -3)
        if "rescale" in item:
            cs0 = cs0 * item["rescale"]
        if "weight_scale" in item:
            cs0 = cs0 * item["weight_scale"]
        item["nu"] = dens * v * _flat_linear(energy0, cs0, energy)      # :63-66
        targets.setdefault(item["target"], []).append(item)
        nprocs += 1

    proc = []
    rate = np.zeros((nprocs + 1, len(energy)))
    i = 0
    for _, ps in targets.items():
        ensure_elastic(ps)
        for item in ps:
            proc.append(_KINDS[item["kind"]](item))
            rate[i] = item["nu"]
            if verbose:
                print("New process: " + signature(item))
            i += 1
    nutotal = rate.sum(axis=0)
    maxrate = float(nutotal.max())
    rate[nprocs] = maxrate - nutotal                            # :114-117
    proc.append(pr.NullCollision())
    inteng = rate.sum(axis=1)
    perm = np.argsort(-inteng, kind="stable")                   # sortperm(inteng, rev=true)
    origperm = np.argsort(perm, kind="stable")                  # invperm
    return {"proc": tuple(proc[k] for k in perm), "rate": np.ascontiguousarray(rate[perm]), "maxrate": maxrate,
            "origperm": origperm}


def lxcat_collision_table(fnames, densities, nE=4096, emax=100.0 * co.eV, emin=None, grid_kind=0, **kw):
    """`load_lxcat` on a LinRange (grid_kind 0, from 0) or LogLinRange (grid_kind 1, from emin) grid, wrapped in the
    CollisionTable the kernels take (collision_table.jl:15-28)."""
    from .tables import loglinrange
    if grid_kind == 0:
        L1, L2 = 0.0, float(emax)
        eng = np.linspace(L1, L2, nE)
    else:
        L1, L2 = math.log(emin if emin is not None else 1e-3 * co.eV), math.log(emax)
        eng = loglinrange(L1, L2, nE)
    r = load_lxcat(fnames, densities, eng, **kw)
    return CollisionTable(proc=list(r["proc"]), grid_kind=grid_kind, L1=L1, L2=L2, nE=nE, rate=r["rate"], maxrate=r["maxrate"])
