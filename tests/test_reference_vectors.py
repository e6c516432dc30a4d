"""Vectors emitted by the REFERENCE itself (julia/emit_golden.jl, run under a Julia runtime with the unmodified Particulator.jl)
against the CPU oracle — the comparison that clears the "parity unpinned" flag of oracle/ptl_oracle.c.

tests/golden/reference_vectors.npz does not exist in this repository: the build image has no Julia.  The test below activates
the moment the file is produced (`julia julia/emit_golden.jl DIR` + `python tests/golden/import_reference_vectors.py DIR`).
Until then `test_comparison_machinery_on_a_synthetic_file` keeps the whole chain exercised — schema, importer, the
injected-uniform replay entry point of the oracle, every comparison — on a file of the same schema written from the oracle."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import particulator_b200 as P
from oracle_backend import oracle_backend, oracle_context

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import import_reference_vectors  # noqa: E402
import make_golden  # noqa: E402

co = P.co
REF = os.path.join(HERE, "golden", "reference_vectors.npz")
SPECIES = {"electron": P.ELECTRON, "positron": P.POSITRON, "photon": P.PHOTON}
KIND_OF = {"NullCollision": 0, "RelativisticCoulomb": 1, "RBEB": 2, "Moller": 3, "Bhaba": 4, "SeltzerBerger": 5, "Compton": 6,
           "PhotoElectric": 7, "BetheHeitler": 8, "PositronAnihilation": 9}


def _air_tables():
    comp = P.air_composition()
    Fdt = co.elementary_charge * 5e5 * 2.5e-11
    return {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}


def _collide_replay(octx, species, tab, j, p3, uniforms):
    """ora_collide_replay: collide() of process j with the given uniforms injected draw by draw (test-only oracle entry point)."""
    dll = oracle_backend().dll
    f = dll.ora_collide_replay
    f.restype = C.c_int32
    dp = C.POINTER(C.c_double)
    f.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, dp, dp, C.c_int32, dp]
    tid = octx.table(tab)
    p3 = np.ascontiguousarray(p3, dtype=np.float64).reshape(-1, 3)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    out = np.zeros((p3.shape[0], 24))
    rc = f(octx.h, species, tid, j, p3.shape[0], p3.ctypes.data_as(dp), u.ctypes.data_as(dp), u.shape[1], out.ctypes.data_as(dp))
    assert rc == 0, rc
    return out


def compare_with_reference(ref, octx, tables, rtol_tables=1e-9, rtol_events=1e-9):
    """Everything emit_golden.jl writes, against the oracle.  Returns the list of array families compared."""
    done = []
    for name, tab in tables.items():
        if f"tables_{name}_rate" in ref:
            kinds = [KIND_OF[type(p).__name__] for p in tab.proc]
            assert ref[f"tables_{name}_prockind"].tolist() == kinds, (name, "process order differs from the reference's")
            # chebfit + every totalcs + compratebound + the sort: coefficient arrays of the builders (row a19)
            np.testing.assert_allclose(tab.rate, ref[f"tables_{name}_rate"], rtol=rtol_tables, atol=1e-6 * np.abs(ref[f"tables_{name}_rate"]).max())
            np.testing.assert_allclose(tab.ratebound, ref[f"tables_{name}_ratebound"], rtol=rtol_tables,
                                       atol=1e-6 * np.abs(ref[f"tables_{name}_ratebound"]).max())
            done.append(f"tables_{name}")
        if f"lookup_{name}_rates" in ref:
            # the LOOKUP arithmetic on the reference's own coefficients: bit-exact tier (rows a5/a6)
            ref_tab = P.ChebyshevCollisionTable(proc=tab.proc, b=tab.b, rate=np.asarray(ref[f"tables_{name}_rate"]),
                                                ratebound=np.asarray(ref[f"tables_{name}_ratebound"]), species=tab.species)
            rates, bound = octx.table_eval(ref_tab, ref[f"lookup_{name}_energy"])
            assert np.array_equal(rates.view(np.uint64), np.asarray(ref[f"lookup_{name}_rates"]).view(np.uint64)), name
            assert np.array_equal(bound.view(np.uint64), np.asarray(ref[f"lookup_{name}_bound"]).view(np.uint64)), name
            done.append(f"lookup_{name}")
        for j in range(len(tab.proc)):
            key = f"collide_{name}_{j}"
            if key + "_out" not in ref:
                continue
            got = _collide_replay(octx, SPECIES[name], tab, j, np.asarray(ref[key + "_p"]).T, np.asarray(ref[key + "_uniforms"]).T)
            want = np.asarray(ref[key + "_out"]).T
            assert np.array_equal(got[:, :3], want[:, :3]), (key, "outcome kinds / species")
            scale = np.maximum(np.abs(want[:, 4:16]).max(axis=1, keepdims=True), 1e-300)
            assert (np.abs(got[:, 4:16] - want[:, 4:16]) / scale).max() <= rtol_events, key
            done.append(key)
    for Z in (7, 8):
        if f"sb_{Z}_data" in ref:
            sb = P.seltzer.from_Z(Z)
            np.testing.assert_allclose(sb.data, ref[f"sb_{Z}_data"], rtol=1e-6, atol=1e-9)       # Newton tolerance of seltzer.jl:193-208
            np.testing.assert_allclose(sb.log_energy, ref[f"sb_{Z}_log_energy"], rtol=1e-14)
            np.testing.assert_allclose([sb.totalcs(k) for k in ref[f"sb_{Z}_K"]], ref[f"sb_{Z}_totalcs"], rtol=1e-9)
            done.append(f"sb_{Z}")
    if "kin_p" in ref:
        p = np.asarray(ref["kin_p"]).T
        np.testing.assert_allclose(P.kinenergy(P.ELECTRON, p), ref["kin_energy"], rtol=1e-14)
        done.append("kin")
    if "push_x" in ref:
        # one free-flight RK2 push through advance!: s = 1e30 keeps the particle collision-free for the whole dt
        p = np.asarray(ref["kin_p"]).T
        n = len(p)
        st = dict(x=np.tile([0.1, -0.2, 0.3], (n, 1)), p=p, s=np.full(n, 1e30))
        pop = P.Population(octx, P.ELECTRON, n + 8, st, tables["electron"], 1e3 * co.eV)
        mp = P.MultiPopulation(("electron", pop))
        psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -5e5]), P.HomogeneousField([0.0, 0.0, 0.0])))
        P.advance(mp, psh, 2.5e-11)
        d = pop.download()
        np.testing.assert_allclose(d["x"], np.asarray(ref["push_x"]).T, rtol=1e-13, atol=1e-16)
        np.testing.assert_allclose(d["p"], np.asarray(ref["push_p"]).T, rtol=1e-13)
        done.append("push")
    case = 1
    while f"repack_{case}_active" in ref:
        act = np.asarray(ref[f"repack_{case}_active"]).astype(np.uint8)
        n = len(act)
        st = dict(x=np.stack([np.arange(1, n + 1, dtype=float), np.zeros(n), np.zeros(n)], axis=1), p=np.tile([0, 0, 1e-21], (n, 1)),
                  active=act)
        pop = P.Population(octx, P.ELECTRON, n + 8, st, tables["electron"], 1e3 * co.eV)
        P.repack(pop)
        assert pop.download()["x"][:, 0].astype(np.int64).tolist() == np.asarray(ref[f"repack_{case}_order"]).tolist(), case
        done.append(f"repack_{case}")
        case += 1
    return done


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/reference_vectors.npz not present: it is written by julia/emit_golden.jl "
                                                    "under a Julia runtime (absent from the build image) — parity stays UNPINNED until then")
def test_oracle_against_vectors_emitted_by_the_reference():
    ref = dict(np.load(REF))
    octx = oracle_context()
    done = compare_with_reference(ref, octx, _air_tables())
    octx.close()
    assert any(k.startswith("collide_") for k in done) and any(k.startswith("lookup_") for k in done)


def _write_synthetic_emit_dir(path, octx, tables):
    """A directory with the layout julia/emit_golden.jl writes, produced from the ORACLE (plumbing test only)."""
    os.makedirs(path, exist_ok=True)
    man = {}

    def emit(name, a):
        a = np.asarray(a)
        dt = {"float64": "Float64", "int64": "Int64", "uint8": "UInt8"}[str(a.dtype)]
        np.asfortranarray(a).ravel(order="F").tofile(os.path.join(path, name + ".bin"))
        man[name] = {"shape": list(a.shape), "dtype": dt, "order": "F"}

    rng = np.random.default_rng(0)
    for name, tab in tables.items():
        emit(f"tables_{name}_rate", tab.rate)
        emit(f"tables_{name}_ratebound", tab.ratebound)
        emit(f"tables_{name}_prockind", np.array([KIND_OF[type(p).__name__] for p in tab.proc], dtype=np.int64))
        e = make_golden.golden_energies(tab.b.xmax)
        rates, bound = octx.table_eval(tab, e)
        emit(f"lookup_{name}_energy", e)
        emit(f"lookup_{name}_rates", np.ascontiguousarray(rates))
        emit(f"lookup_{name}_bound", bound)
        for j, proc in enumerate(tab.proc):
            lo = make_golden.LOW[name]
            if proc.name == "BetheHeitler":
                lo = 1.05e6
            if proc.name == "RBEB":
                lo = max(lo, 1.05 * proc.B / co.eV)
            p3 = make_golden.golden_momenta(SPECIES[name], lo)
            u = rng.random((len(p3), 64))
            out = _collide_replay(octx, SPECIES[name], tab, j, p3, u)
            out[:, 3] = 0
            emit(f"collide_{name}_{j}_p", p3.T)
            emit(f"collide_{name}_{j}_uniforms", u.T)
            emit(f"collide_{name}_{j}_out", out.T)
    p = make_golden.golden_momenta(P.ELECTRON, 1.2e3, n=32)
    emit("kin_p", p.T)
    emit("kin_energy", P.kinenergy(P.ELECTRON, p))
    for case, (n, frac) in enumerate([(1, 0.0), (31, 0.5), (1025, 0.9)], start=1):
        act = (np.random.default_rng(n).random(n) >= frac).astype(np.uint8)
        st = dict(x=np.stack([np.arange(1, n + 1, dtype=float), np.zeros(n), np.zeros(n)], axis=1), p=np.tile([0, 0, 1e-21], (n, 1)), active=act)
        pop = P.Population(octx, P.ELECTRON, n + 8, st, tables["electron"], 1e3 * co.eV)
        P.repack(pop)
        emit(f"repack_{case}_active", act)
        emit(f"repack_{case}_order", pop.download()["x"][:, 0].astype(np.int64))
    json.dump({"format": "particulator_b200.reference_vectors", "version": 1, "julia": "synthetic (oracle)", "arrays": man},
              open(os.path.join(path, "manifest.json"), "w"))


def test_comparison_machinery_on_a_synthetic_file(tmp_path):
    tables = _air_tables()
    octx = oracle_context()
    _write_synthetic_emit_dir(str(tmp_path / "emit"), octx, tables)
    dst, count = import_reference_vectors.pack(str(tmp_path / "emit"), str(tmp_path / "ref.npz"))
    ref = dict(np.load(dst))
    done = compare_with_reference(ref, octx, tables)
    octx.close()
    assert count > 60 and len(done) > 30
    assert {"tables_electron", "lookup_photon", "kin", "repack_3"} <= set(done)
