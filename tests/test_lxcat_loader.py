"""Host-init builder `load_lxcat` (reference src/lxcat.jl:17-134) against a literal, independent
recomputation of every rule it applies, then the loaded table through the oracle's sub-step loop."""
import numpy as np
import pytest

import particulator_b200 as P
from particulator_b200 import lxcat
import lxcat_fixture

co = P.co


@pytest.fixture()
def dbfile(tmp_path):
    return lxcat_fixture.write(str(tmp_path / "db" / "air.json"))


def _interp_flat(x0, y0, x):
    out = np.empty_like(x)
    for i, xv in enumerate(x):          # scalar loop on purpose: independent of the vectorised implementation
        if xv <= x0[0]:
            out[i] = y0[0]
        elif xv >= x0[-1]:
            out[i] = y0[-1]
        else:
            k = max(j for j in range(len(x0)) if x0[j] <= xv)
            f = (xv - x0[k]) / (x0[k + 1] - x0[k])
            out[i] = (1 - f) * y0[k] + f * y0[k + 1]
    return out


def test_load_lxcat_rules(dbfile):
    dens = {"N2": 0.79 * co.nair, "O2": 0.21 * co.nair}
    eng = np.linspace(0.0, 120.0 * co.eV, 257)       # beyond the last knot: flat extrapolation
    r = lxcat.load_lxcat([dbfile], dens, eng)
    recs = lxcat_fixture.records()
    v = np.sqrt(2 * eng / co.electron_mass)

    def nu(rec, scale=1.0):
        x0 = np.array([d[0] for d in rec["data"]]) * co.eV
        y0 = np.array([d[1] for d in rec["data"]]) * scale
        return dens[rec["target"]] * v * _interp_flat(x0, y0, eng)
    exp = {
        "N2 exc": nu(recs[1], 0.8), "N2 ion": nu(recs[2]),
        "O2 el": nu(recs[3]), "O2 att": nu(recs[4], 2.0 * dens["O2"] / 1e6), "O2 exc": nu(recs[5]),
    }
    exp["N2 el"] = nu(recs[0]) - exp["N2 exc"] - exp["N2 ion"]        # ensure_elastic
    rows = [exp["N2 el"], exp["N2 exc"], exp["N2 ion"], exp["O2 el"], exp["O2 att"], exp["O2 exc"]]
    tot = np.sum(rows, axis=0)
    rows.append(tot.max() - tot)
    rows = np.array(rows)
    assert len(r["proc"]) == 7 and r["rate"].shape == (7, 257)        # Ar skipped (density 0)
    assert r["maxrate"] == pytest.approx(tot.max(), rel=1e-14)
    # rows come back sorted by descending summed rate, origperm maps the pre-sort order to it
    s = r["rate"].sum(axis=1)
    assert np.all(np.diff(s) <= 0)
    np.testing.assert_allclose(r["rate"][r["origperm"]], rows, rtol=1e-12, atol=1e-3)
    kinds = [type(r["proc"][k]).__name__ for k in r["origperm"]]
    assert kinds == ["Elastic", "Excitation", "Ionization", "Elastic", "Attachment", "Excitation", "NullCollision"]
    pr = [r["proc"][k] for k in r["origperm"]]
    assert pr[0].mass_ratio == 1.95e-5 and pr[2].threshold == pytest.approx(15.6 * co.eV)
    # every column sums to maxrate: the explicit null row closes the budget (lxcat.jl:114-117)
    np.testing.assert_allclose(r["rate"].sum(axis=0), r["maxrate"], rtol=1e-12)
    assert np.all(r["rate"] >= -1e-3)


def test_photoemission_is_refused(tmp_path):
    import json
    p = tmp_path / "pe.json"
    p.write_text(json.dumps([{"target": "N2", "kind": "PHOTOEMISSION", "comment": "", "data": [[0, 1e-22], [10, 1e-22]]}]))
    with pytest.raises(NotImplementedError):
        lxcat.load_lxcat(str(p), {"N2": 1e25}, np.linspace(0, 1e-18, 8))


def test_loaded_table_runs_through_the_oracle(dbfile, octx):
    dens = {"N2": 0.79 * co.nair, "O2": 0.21 * co.nair}
    tab = lxcat.lxcat_collision_table(dbfile, dens, nE=2048, emax=100 * co.eV)
    assert tab.rate.shape == (7, 2048)
    n = 2000
    rng = np.random.default_rng(5)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * co.eV / co.electron_mass) * 1.2,
              s=-np.log(1 - rng.random(n)), uid=np.arange(1, n + 1, dtype=np.uint64))
    E = 100 * co.Td * co.nair
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -E]), None))
    octx.set_rng(1, 0)
    pop = P.Population(octx, P.SLOW_ELECTRON, 4 * n, st, tab, 0.0)
    mp = P.MultiPopulation(("slow", pop))
    t = 0.0
    for _ in range(5):
        t += 1e-12
        P.advance(mp, psh, t)
    stt = P.last_advance_stats(mp)
    # null-collision stepping: sub-steps per particle-step = maxrate*dt + 1 on average
    kappa = stt["substeps"] / stt["rows"]
    assert kappa == pytest.approx(tab.maxrate * 1e-12 + 1, rel=0.05)
    d = pop.download()
    assert np.all(np.isfinite(d["p"])) and np.all(np.abs(d["t"][d["active"] != 0] - t) <= 2.3e-16)   # loop ends at trem <= eps(Float64) s (mixed_population.jl:66)
