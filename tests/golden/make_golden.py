"""Generate tests/golden/oracle_vectors.npz — regression vectors of the CPU ORACLE (not of the reference: the
reference ships no golden vectors and Julia cannot run in this image).

The file pins, for fixed seeds: table lookups of the three air tables, one collide() event per process for a set of
fixed momenta, and the per-particle end state of a small mixed-population advance!.  Both the oracle (CPU suite) and the
CUDA path (GPU suite) are compared against it, so that neither side can drift silently.  The air tables are built with
the SYNTHETIC Seltzer-Berger stand-in so that the vectors do not depend on the Geant4 data files.

Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import particulator_b200 as P  # noqa: E402
from particulator_b200 import seltzer  # noqa: E402

co = P.co
DT = 2.5e-11


def golden_tables():
    comp = P.air_composition()
    Fdt = co.elementary_charge * 5e5 * DT
    sb = {Z: seltzer.build_from_raw(seltzer._SyntheticRaw(Z), Z, ncum=200, synthetic=True) for Z in (7, 8)}
    return {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15, sb=sb),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}


def golden_energies(xmax):
    return np.concatenate([np.exp(np.linspace(np.log(1e-2 * co.eV), np.log(0.99 * xmax), 97)), [1e3 * co.eV, 0.0]])


def golden_momenta(species, lo_eV, n=16):
    K = np.exp(np.linspace(np.log(lo_eV), np.log(1.5e8), n)) * co.eV
    pn = P.momentum_norm_from_kin(species, K)
    ang = np.linspace(0.3, 2.7, n)
    d = np.stack([np.sin(ang) * np.cos(2.1 * ang), np.sin(ang) * np.sin(2.1 * ang), np.cos(ang)], axis=1)
    return d * pn[:, None]


SPECIES = {"electron": P.ELECTRON, "positron": P.POSITRON, "photon": P.PHOTON}
LOW = {"electron": 1.2e3, "positron": 3e2, "photon": 1.2e3}


def compute(ctx):
    """Everything the golden file holds, computed through `ctx` (oracle or CUDA)."""
    from conftest import make_world, default_pusher
    tabs = golden_tables()
    out = {}
    for name, tab in tabs.items():
        e = golden_energies(tab.b.xmax)
        rates, bound = ctx.table_eval(tab, e)
        out[f"lookup_{name}_rates"] = np.ascontiguousarray(rates)
        out[f"lookup_{name}_bound"] = bound
        ctx.set_rng(2024, 7)
        for j, proc in enumerate(tab.proc):
            lo = LOW[name]
            if proc.name == "BetheHeitler":
                lo = 1.05e6
            if proc.name == "RBEB":
                lo = max(lo, 1.05 * proc.B / co.eV)
            out[f"collide_{name}_{j}"] = ctx.collide_test(SPECIES[name], tab, j, golden_momenta(SPECIES[name], lo), uid0=4242)[:, :16]
    ctx.set_rng(11, 0)
    mp, el, ph, po = make_world(ctx, tabs, 300, 300, 100, cap=6000, seed=5)
    P.advance(mp, default_pusher(), DT)
    for nm, q in (("electron", el), ("photon", ph), ("positron", po)):
        d = q.download()
        o = np.argsort(d["uid"], kind="stable")
        for k in ("x", "p", "t", "s", "r", "active", "uid"):
            out[f"advance_{nm}_{k}"] = d[k][o]
    st = P.last_advance_stats(mp)
    out["advance_substeps"] = np.array([st["substeps"]])
    return out


if __name__ == "__main__":
    from oracle_backend import oracle_context
    res = compute(oracle_context())
    path = os.path.join(HERE, "oracle_vectors.npz")
    np.savez_compressed(path, **res)
    print(path, os.path.getsize(path), "bytes,", len(res), "arrays")
