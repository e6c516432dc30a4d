"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tiers:
  * bit-exact: RNG stream, table lookups (Chebyshev and LinRange), droplow!/repack! permutation;
  * deterministic replay: same Philox deviates through both codes, per-particle state matched by uid,
    rel 1e-6 in fp64 (tolerance written below as REPLAY_RTOL);
  * properties at larger sizes that do not need the oracle.
Run on a B200 with:  python -m pytest tests -m gpu -x -q"""
import json
import os

import numpy as np
import pytest

import particulator_b200 as P
from conftest import make_world, default_pusher

pytestmark = pytest.mark.gpu

co = P.co
REPLAY_RTOL = 1e-6     # north-star bound for the deterministic path in fp64
EVENT_RTOL = 1e-9      # single collide() events (no error accumulation)
DT = 2.5e-11


# ---------------------------------------------------------------------------------------------------
# bit-exact tier
# ---------------------------------------------------------------------------------------------------
def test_rng_stream_bit_exact(gctx, octx):
    for uid, seed, step in [(1, 0, 0), (2 ** 40 + 17, 12345678901234567, 3), (2 ** 64 - 1, 2 ** 64 - 1, 2 ** 32 - 1)]:
        a = gctx.rng_test(uid, seed, step, 257)
        b = octx.rng_test(uid, seed, step, 257)
        assert np.array_equal(a, b)
        assert np.all((a > 0) & (a < 1))


def _energies(xmax):
    rng = np.random.default_rng(1)
    e = np.exp(rng.uniform(np.log(1e-3 * co.eV), np.log(0.9999 * xmax), 200000))
    edges = xmax * 2.0 ** np.arange(-40, 0)
    edge_lo = np.nextafter(edges, 0)
    edge_hi = np.nextafter(edges, np.inf)
    return np.concatenate([e, edges, edge_lo, edge_hi, [0.0, 1e3 * co.eV, 1e2 * co.eV]])


@pytest.mark.parametrize("name", ["electron", "photon", "positron"])
def test_chebyshev_table_lookup_bit_exact(gctx, octx, air_tables, name):
    tab = air_tables[name]
    e = _energies(tab.b.xmax)
    rg, bg = gctx.table_eval(tab, e)
    ro, bo = octx.table_eval(tab, e)
    assert np.array_equal(rg.view(np.uint64), ro.view(np.uint64))
    assert np.array_equal(bg.view(np.uint64), bo.view(np.uint64))
    assert gctx.error_flags(clear=True) == 0


def test_linear_table_lookup(gctx, octx):
    lin = P.synthetic_lxcat_table(grid_kind=0)
    e = np.random.default_rng(2).uniform(0, 99.99, 100000) * co.eV
    rg, bg = gctx.table_eval(lin, e)
    ro, bo = octx.table_eval(lin, e)
    assert np.array_equal(rg.view(np.uint64), ro.view(np.uint64))      # LinRange grid: bit-exact
    assert np.array_equal(bg, bo)
    loglin = P.synthetic_lxcat_table(grid_kind=1)
    e = np.exp(np.random.default_rng(3).uniform(np.log(2e-3), np.log(99.0), 100000)) * co.eV
    rg, _ = gctx.table_eval(loglin, e)
    ro, _ = octx.table_eval(loglin, e)
    # LogLinRange needs log/exp (util.jl:118-127): libm vs libdevice differ in the last ulp
    np.testing.assert_allclose(rg, ro, rtol=1e-9, atol=1e-9 * loglin.maxrate)


def _random_pop(rng, n, frac_dead):
    x = rng.normal(size=(n, 3))
    K = np.exp(rng.uniform(np.log(5e2), np.log(1e7), n)) * co.eV
    pn = P.momentum_norm_from_kin(P.ELECTRON, K)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return dict(x=x, p=d * pn[:, None], w=rng.uniform(0.5, 2, n), t=rng.uniform(0, 1e-9, n), s=rng.uniform(0, 3, n),
                r=rng.uniform(0, 1e12, n), active=(rng.random(n) >= frac_dead).astype(np.uint8),
                uid=np.arange(1, n + 1, dtype=np.uint64))


@pytest.mark.parametrize("n,frac_dead", [(1, 0.0), (1, 1.0), (31, 0.5), (1024, 0.3), (1025, 0.9), (100003, 0.1), (100003, 0.999),
                                          (5000, 0.0), (5000, 1.0), (262144, 0.5)])
def test_droplow_repack_same_permutation(gctx, octx, air_tables, n, frac_dead):
    rng = np.random.default_rng(n)
    st = _random_pop(rng, n, frac_dead)
    res = []
    for ctx in (gctx, octx):
        pop = P.Population(ctx, P.ELECTRON, n + 10, st, air_tables["electron"], 1e3 * co.eV)
        n1 = P.repack(pop)
        a = pop.download()
        n2 = P.droplow(pop)
        b = pop.download()
        res.append((n1, a, n2, b))
    (n1g, ag, n2g, bg), (n1o, ao, n2o, bo) = res
    assert n1g == n1o == int(st["active"].sum())
    assert n2g == n2o
    for g, o in ((ag, ao), (bg, bo)):
        for k in g:
            assert np.array_equal(g[k], o[k]), k       # same rows in the same order, bit for bit


def test_diagnostics_match_oracle(gctx, octx, air_tables):
    st = _random_pop(np.random.default_rng(5), 77777, 0.2)
    d = []
    for ctx in (gctx, octx):
        pop = P.Population(ctx, P.ELECTRON, 80000, st, air_tables["electron"], 1e3 * co.eV)
        d.append((pop.diag(), P.meanenergy(pop), P.spread(pop), P.posvar(pop), P.maxenergy(pop), P.nactives(pop), P.weight(pop)))
    (dg, meg, spg, pvg, mxg, nag, wg), (do, meo, spo, pvo, mxo, nao, wo) = d
    assert nag == nao and dg.n == do.n
    assert mxg == pytest.approx(mxo, rel=1e-14)     # kinenergy: FMA contraction on the device, none in the oracle
    np.testing.assert_allclose([wg, meg], [wo, meo], rtol=1e-12)
    np.testing.assert_allclose(spg[0], spo[0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(spg[1], spo[1], rtol=1e-9)
    np.testing.assert_allclose(pvg, pvo, rtol=1e-9)
    hg = gctx and P.Population(gctx, P.ELECTRON, 80000, st, air_tables["electron"], 1e3 * co.eV).histogram("energy", 1e3 * co.eV, 1e7 * co.eV, 64, True)
    ho = P.Population(octx, P.ELECTRON, 80000, st, air_tables["electron"], 1e3 * co.eV).histogram("energy", 1e3 * co.eV, 1e7 * co.eV, 64, True)
    np.testing.assert_allclose(hg, ho, rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------------------------------------------
# deterministic replay: single collide() events, every process
# ---------------------------------------------------------------------------------------------------
def _momenta(species, n, lo_eV, hi_eV, seed):
    rng = np.random.default_rng(seed)
    K = np.exp(rng.uniform(np.log(lo_eV), np.log(hi_eV), n)) * co.eV
    pn = P.momentum_norm_from_kin(species, K)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return d * pn[:, None]


CASES = [("electron", P.ELECTRON, 1.1e3, 2e8), ("positron", P.POSITRON, 2.5e2, 2e8), ("photon", P.PHOTON, 1.1e3, 2e8)]


@pytest.mark.parametrize("name,species,lo,hi", CASES)
def test_collide_events_replay(gctx, octx, air_tables, name, species, lo, hi):
    tab = air_tables[name]
    gctx.set_rng(7, 3)
    octx.set_rng(7, 3)
    for j, proc in enumerate(tab.proc):
        lo_j = lo
        if proc.name == "BetheHeitler":
            lo_j = 1.03e6          # above 2 mc^2
        if proc.name == "RBEB":
            lo_j = max(lo, 1.05 * proc.B / co.eV)
        p3 = _momenta(species, 20000, lo_j, hi, 100 + j)
        g = gctx.collide_test(species, tab, j, p3, uid0=1000)
        o = octx.collide_test(species, tab, j, p3, uid0=1000)
        same_flow = (g[:, 3] == o[:, 3]) & (g[:, 0] == o[:, 0])
        # control-flow flips of a rejection test are possible only when a compare is within rounding of its
        # threshold: allow at most 2 in 20000 events and require agreement everywhere else
        assert (~same_flow).sum() <= 2, (proc.name, int((~same_flow).sum()))
        assert np.array_equal(g[same_flow, :3], o[same_flow, :3]), proc.name
        scale = np.linalg.norm(p3, axis=1)[same_flow, None]
        for c0 in (4, 8, 12):
            err = np.abs(g[same_flow, c0:c0 + 3] - o[same_flow, c0:c0 + 3]) / scale
            assert err.max() <= EVENT_RTOL, (proc.name, c0, err.max())
        np.testing.assert_allclose(g[same_flow][:, [7, 11, 15]], o[same_flow][:, [7, 11, 15]], rtol=1e-12, atol=0)
    assert gctx.error_flags(clear=True) == 0 and octx.error_flags(clear=True) == 0


def test_lxcat_collide_events_replay(gctx, octx):
    tab = P.synthetic_lxcat_table()
    gctx.set_rng(1, 1)
    octx.set_rng(1, 1)
    rng = np.random.default_rng(4)
    v = rng.normal(size=(5000, 3)) * 1.5e6
    for j, proc in enumerate(tab.proc):
        g = gctx.collide_test(P.SLOW_ELECTRON, tab, j, v, uid0=5)
        o = octx.collide_test(P.SLOW_ELECTRON, tab, j, v, uid0=5)
        assert np.array_equal(g[:, :4], o[:, :4]), proc.name
        np.testing.assert_allclose(g[:, 4:16], o[:, 4:16], rtol=1e-9, atol=1e-3)


# ---------------------------------------------------------------------------------------------------
# deterministic replay: full advance! steps, per-particle state matched by uid
# ---------------------------------------------------------------------------------------------------
def _by_uid(pop):
    d = pop.download()
    order = np.argsort(d["uid"], kind="stable")
    return {k: v[order] for k, v in d.items()}


#: observed per-test mismatch counts (label -> (mismatching particles, particles compared)); written to
#: gpurun_out/replay_mismatches.json at the end of the session and asserted to be zero test by test
OBSERVED_MISMATCHES = {}
#: PTL_REPLAY_BUDGET=<fraction> relaxes the zero-mismatch assertion (the count is still printed and recorded)
_BUDGET_OVERRIDE = float(os.environ.get("PTL_REPLAY_BUDGET", "0"))


@pytest.fixture(scope="session", autouse=True)
def _record_mismatches():
    yield
    if OBSERVED_MISMATCHES:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        try:
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "replay_mismatches.json"), "w") as f:
                json.dump({"total_mismatches": sum(v[0] for v in OBSERVED_MISMATCHES.values()),
                           "total_compared": sum(v[1] for v in OBSERVED_MISMATCHES.values()), "by_test": OBSERVED_MISMATCHES}, f, indent=1)
        except OSError:
            pass


def _compare_populations(gp, op, label, budget=0.0):
    """Per-particle comparison matched by uid.  A mismatching particle is one whose trajectory took a different branch at a
    compare within rounding of its threshold.  Round 1 allowed 0.2 % of them silently; the observed count has always been 0,
    so it is now ASSERTED to be zero (budget = 0), printed, and recorded."""
    budget = max(budget, _BUDGET_OVERRIDE)
    g, o = _by_uid(gp), _by_uid(op)
    ug, uo = g["uid"], o["uid"]
    common, ig, io = np.intersect1d(ug, uo, return_indices=True)
    nmiss = (len(ug) - len(common)) + (len(uo) - len(common))
    assert nmiss <= budget * max(len(ug), 1), (label, "uids present on one side only", nmiss, len(ug), len(uo), len(common))
    if len(common) == 0:
        OBSERVED_MISMATCHES[label] = (nmiss, 0)
        return 0
    pscale = np.maximum(np.linalg.norm(o["p"][io], axis=1), 1e-300)[:, None]
    bad = (np.abs(g["p"][ig] - o["p"][io]) / pscale).max(axis=1) > REPLAY_RTOL
    bad |= (np.abs(g["x"][ig] - o["x"][io])).max(axis=1) > REPLAY_RTOL * np.maximum(np.abs(o["x"][io]).max(axis=1), co.c * DT)
    bad |= np.abs(g["t"][ig] - o["t"][io]) > 1e-6 * DT
    bad |= np.abs(g["s"][ig] - o["s"][io]) > REPLAY_RTOL * np.maximum(np.abs(o["s"][io]), 1.0)
    bad |= np.abs(g["r"][ig] - o["r"][io]) > REPLAY_RTOL * np.maximum(np.abs(o["r"][io]), 1.0)
    bad |= g["active"][ig] != o["active"][io]
    bad |= g["w"][ig] != o["w"][io]
    nbad = int(bad.sum()) + nmiss
    OBSERVED_MISMATCHES[label] = (nbad, int(len(common)))
    print(f"[replay] {label}: {nbad} mismatching of {len(common)} particles compared")
    assert bad.sum() <= budget * len(common), (label, "mismatching particles", int(bad.sum()), len(common))
    return nbad


@pytest.mark.parametrize("seed,n_e,n_g,n_p", [(0, 3000, 3000, 1000), (1, 257, 0, 0), (2, 0, 5000, 0), (3, 0, 0, 777)])
def test_advance_replay_against_oracle(gctx, octx, air_tables, seed, n_e, n_g, n_p):
    worlds = []
    for ctx in (gctx, octx):
        ctx.set_rng(seed, 0)
        worlds.append(make_world(ctx, air_tables, n_e, n_g, n_p, cap=40000, seed=seed))
    psh = default_pusher()
    t = 0.0
    for step in range(2):
        t += DT
        for mp, *_ in worlds:
            P.advance(mp, psh, t)
        sg, so = P.last_advance_stats(worlds[0][0]), P.last_advance_stats(worlds[1][0])
        assert abs(sg["substeps"] - so["substeps"]) <= 2e-3 * so["substeps"] + 50, (sg, so)
        for k, label in ((1, "electron"), (2, "photon"), (3, "positron")):
            _compare_populations(worlds[0][k], worlds[1][k], f"{label} step {step}")
        for mp, *pops in worlds:
            for q in pops:
                P.droplow(q)
        for k in (1, 2, 3):
            assert abs(len(worlds[0][k]) - len(worlds[1][k])) <= 3


def test_below_cut_coasting_matches_substep_loop(gctx, octx, air_tables):
    """Electrons that the field decelerates below energy_cut keep r != 0 and repeat the same s/r sub-step without
    drawing (collisions.jl:148-151, mixed_population.jl:66-87).  The CUDA path takes those sub-steps in blocks
    (wf_coast_below_cut); the oracle takes them one by one.  Same sub-step count, same t/s/r, x and p within the
    replay tolerance -- including particles whose energy comes back above the cut inside the step."""
    n = 4000
    rng = np.random.default_rng(77)
    K = (1e3 * (1 + 10 ** rng.uniform(-5, -1.3, n))) * co.eV              # just above the 1 keV cut
    pn = P.momentum_norm_from_kin(P.ELECTRON, K)
    cost = np.where(rng.random(n) < 0.5, rng.uniform(-1, -0.2, n), rng.uniform(-0.08, 0.0, n))   # against the force / nearly transverse
    phi = rng.uniform(0, 2 * np.pi, n)
    sint = np.sqrt(1 - cost ** 2)
    d = np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)
    st = dict(x=rng.normal(0, 1.0, (n, 3)), p=d * pn[:, None], s=10 ** rng.uniform(-3.5, 0.3, n),
              uid=np.arange(1, n + 1, dtype=np.uint64))
    worlds = []
    for ctx in (gctx, octx):
        ctx.set_rng(5, 0)
        el = P.Population(ctx, P.ELECTRON, 4 * n, st, air_tables["electron"], 1e3 * co.eV)
        ph = P.Population(ctx, P.PHOTON, 4 * n, None, air_tables["photon"], 1e3 * co.eV)
        worlds.append((P.MultiPopulation(("electron", el), ("photon", ph)), el, ph))
    psh = default_pusher()
    for mp, *_ in worlds:
        P.advance(mp, psh, DT)
    sg, so = P.last_advance_stats(worlds[0][0]), P.last_advance_stats(worlds[1][0])
    assert so["substeps"] > 100 * n                     # the case really is dominated by the repeated sub-steps
    assert abs(sg["substeps"] - so["substeps"]) <= 1e-3 * so["substeps"], (sg, so)
    _compare_populations(worlds[0][1], worlds[1][1], "below-cut electrons")
    below = P.kinenergy(P.ELECTRON, worlds[1][1].download()["p"]) < 1e3 * co.eV
    assert 0.2 < below.mean() < 0.999                   # both outcomes are present: still below, and back above the cut


def test_lepton_streaming_path_replay(gctx, octx, air_tables):
    """Leptons at kappa ~ 1 (dt scaled down 2048x): from the second advance! on the library routes the species through the
    streaming kernels (TMA-staged k_advance_stream_tma by default, k_advance_stream with the option off) and defers only the
    rows that collide within dt.  Both must replay the oracle particle by particle and agree with each other bit for bit;
    6001 rows: a ragged last tile (not a multiple of the 256-row tile nor of the 16-byte copy granule)."""
    dt = DT / 2048
    results = []
    for tma in (1, 0):
        gctx.set_option("stream_tma", tma)
        worlds = []
        for ctx in (gctx, octx) if tma else (gctx,):
            ctx.set_rng(11, 0)
            worlds.append(make_world(ctx, air_tables, 6001, 0, 0, cap=20000, seed=21))
        psh = default_pusher()
        t = 0.0
        for step in range(4):
            t += dt
            for mp, *_ in worlds:
                P.advance(mp, psh, t)
            if tma:
                sg, so = P.last_advance_stats(worlds[0][0]), P.last_advance_stats(worlds[1][0])
                assert sg["substeps"] == so["substeps"], (step, sg, so)
                assert so["substeps"] < 1.5 * 6001              # the case really is kappa ~ 1
                _compare_populations(worlds[0][1], worlds[1][1], f"streaming electrons step {step}")
        results.append(_by_uid(worlds[0][1]))
    gctx.set_option("stream_tma", 1)
    for k in results[0]:
        assert np.array_equal(results[0][k], results[1][k]), k     # the two streaming kernels: identical bits
    assert gctx.error_flags(clear=True) == 0


def test_photon_free_flight_and_time(gctx, air_tables):
    """Photons (kappa ~ 1e-4): x advances by c*dt along p, t == tfinal, p untouched for non-colliding ones."""
    mp, el, ph, po = make_world(gctx, air_tables, 0, 200000, 0, cap=300000, seed=9)
    before = ph.download()
    P.advance(mp, default_pusher(), DT)
    after = ph.download()
    n = len(before["w"])
    same = np.all(after["p"][:n] == before["p"], axis=1) & (after["active"][:n] == 1)
    assert same.mean() > 0.97
    d = before["p"][same] / np.linalg.norm(before["p"][same], axis=1)[:, None]
    np.testing.assert_allclose(after["x"][:n][same], before["x"][same] + d * co.c * DT, rtol=0, atol=1e-12)
    np.testing.assert_allclose(after["t"][:n][after["active"][:n] == 1], DT, rtol=0, atol=2.3e-16)


def test_wall_and_counter_callbacks(gctx, octx, air_tables):
    res = []
    for ctx in (gctx, octx):
        ctx.set_rng(11, 0)
        mp, el, ph, po = make_world(ctx, air_tables, 2000, 2000, 0, cap=20000, seed=4, emin=1e5, emax=1e7)
        wall_e = P.WallCallback(P.ELECTRON, 3, 0.001, drop=True)
        wall_g = P.WallCallback(P.PHOTON, 3, 0.002, drop=False)
        cc = P.CollisionCounter()
        cb = P.CombinedCallback((wall_e, wall_g, cc))
        P.advance(mp, default_pusher(), DT, cb)
        res.append((wall_e.accum, wall_g.accum, dict(cc.d), P.nactives(el), P.nactives(ph)))
    (weg, wgg, cg, nag, npg), (weo, wgo, co_, nao, npo) = res
    assert len(weg["w"]) == len(weo["w"]) > 0
    assert len(wgg["w"]) == len(wgo["w"]) > 0
    assert nag == nao and npg == npo
    for a, b in ((weg, weo), (wgg, wgo)):
        ka, kb = np.lexsort(a["x"].T), np.lexsort(b["x"].T)
        np.testing.assert_allclose(a["x"][ka], b["x"][kb], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(a["p"][ka], b["p"][kb], rtol=1e-6, atol=1e-30)
    assert set(cg) == set(co_)
    for k in cg:
        assert abs(cg[k] - co_[k]) <= 2 + 2e-3 * co_[k], (k, cg[k], co_[k])


def test_capacity_overflow_is_flagged(gctx, air_tables):
    mp, el, ph, po = make_world(gctx, air_tables, 3000, 0, 0, cap=3005, seed=6, emin=1e6, emax=1e7)
    rc = P.advance(mp, default_pusher(), DT, check=False)
    assert rc & 1                      # PTL_ERR_CAPACITY_OVERFLOW  (population.jl:107)
    assert len(el) <= 3005
    assert gctx.error_flags(clear=True) & 1


@pytest.mark.parametrize("extra_levels", [0, 57])
def test_slow_electron_replay(gctx, octx, extra_levels):
    """Config 5: LXCat linear table, null-collision dominated, attachment deaths and ionisation births.  With 57 extra
    excitation channels (64 rows, the size of a real N2/O2 set) the kernels select by binary search over the running sums."""
    tab = P.synthetic_lxcat_table(extra_levels=extra_levels)
    n = 5000
    rng = np.random.default_rng(8)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * co.eV / co.electron_mass) * 1.2,
              s=-np.log(1 - rng.random(n)), uid=np.arange(1, n + 1, dtype=np.uint64))
    E = 100 * co.Td * co.nair
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -E]), None))
    pops = []
    for ctx in (gctx, octx):
        ctx.set_rng(3, 0)
        pop = P.Population(ctx, P.SLOW_ELECTRON, 4 * n, st, tab, 0.0)
        mp = P.MultiPopulation(("slow", pop))
        t = 0.0
        for _ in range(3):
            t += 1e-12
            P.advance(mp, psh, t)
        pops.append((pop, P.last_advance_stats(mp)))
    (pg, sg), (po_, so) = pops
    assert abs(sg["substeps"] - so["substeps"]) <= 2e-3 * so["substeps"] + 20
    g, o = _by_uid(pg), _by_uid(po_)
    common, ig, io = np.intersect1d(g["uid"], o["uid"], return_indices=True)
    assert len(common) >= 0.998 * max(len(g["uid"]), len(o["uid"]))
    vs = np.maximum(np.linalg.norm(o["p"][io], axis=1), 1e3)[:, None]
    bad = (np.abs(g["p"][ig] - o["p"][io]) / vs).max(axis=1) > REPLAY_RTOL
    bad |= g["active"][ig] != o["active"][io]
    assert bad.sum() <= max(2, 0.002 * len(common)), int(bad.sum())


def test_lxcat_json_table_replay(gctx, octx, tmp_path):
    """A table built by `load_lxcat` from a JSON database (EFFECTIVE->ELASTIC, 3-body attachment, rescale) drives the
    same kernels: deterministic replay against the oracle."""
    import lxcat_fixture
    f = lxcat_fixture.write(str(tmp_path / "air.json"))
    tab = P.lxcat_collision_table(f, {"N2": 0.79 * co.nair, "O2": 0.21 * co.nair}, nE=2048, emax=100 * co.eV)
    n = 4000
    rng = np.random.default_rng(11)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * co.eV / co.electron_mass) * 1.5,
              s=-np.log(1 - rng.random(n)), uid=np.arange(1, n + 1, dtype=np.uint64))
    E = 150 * co.Td * co.nair
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -E]), None))
    pops = []
    for ctx in (gctx, octx):
        ctx.set_rng(5, 0)
        pop = P.Population(ctx, P.SLOW_ELECTRON, 4 * n, st, tab, 0.0)
        mp = P.MultiPopulation(("slow", pop))
        t = 0.0
        for _ in range(3):
            t += 1e-12
            P.advance(mp, psh, t)
        pops.append((pop, P.last_advance_stats(mp)))
    (pg, sg), (po_, so) = pops
    assert abs(sg["substeps"] - so["substeps"]) <= 2e-3 * so["substeps"] + 20
    g, o = _by_uid(pg), _by_uid(po_)
    common, ig, io = np.intersect1d(g["uid"], o["uid"], return_indices=True)
    assert len(common) >= 0.998 * max(len(g["uid"]), len(o["uid"]))
    vs = np.maximum(np.linalg.norm(o["p"][io], axis=1), 1e3)[:, None]
    bad = (np.abs(g["p"][ig] - o["p"][io]) / vs).max(axis=1) > REPLAY_RTOL
    bad |= g["active"][ig] != o["active"][io]
    assert bad.sum() <= max(2, 0.002 * len(common)), int(bad.sum())


def test_context_driven_from_another_host_thread(gctx, air_tables):
    """One host thread per context is the contract, but not necessarily the thread that created it (bench.py's e2e leg and a
    Julia task pool both hand contexts to worker threads): every entry point binds the caller to the context's device."""
    import threading
    mp, el, ph, po = make_world(gctx, air_tables, 3000, 500, 50, cap=30000, seed=2)
    res = {}

    def work():
        try:
            P.advance(mp, default_pusher(), DT)
            for q in (el, ph, po):
                P.droplow(q)
            res["n"] = len(el)
            res["p"] = el.download()["p"]
        except Exception as exc:      # pragma: no cover
            res["err"] = repr(exc)

    th = threading.Thread(target=work)
    th.start(); th.join()
    assert "err" not in res, res.get("err")
    assert res["n"] > 0 and np.all(np.isfinite(res["p"]))


# ---------------------------------------------------------------------------------------------------
# size-independent properties at larger sizes (no oracle)
# ---------------------------------------------------------------------------------------------------
def test_large_population_properties(gctx, air_tables):
    n = 2_000_000
    mp, el, ph, po = make_world(gctx, air_tables, n, 0, 0, cap=int(1.3 * n), seed=10, emin=1e4, emax=5e7)
    w0 = P.weight(el)
    P.advance(mp, default_pusher(), DT)
    st = P.last_advance_stats(mp)
    assert st["substeps"] > 30 * n                  # kappa > 30 in STP air at dt = 2.5e-11 s
    d = el.download(("t", "active", "uid"))
    # the loop stops when trem <= eps(Float64) = 2.2e-16 *seconds* (absolute, mixed_population.jl:66)
    assert np.all(np.abs(d["t"][d["active"] == 1] - DT) <= 2.3e-16)
    assert len(np.unique(d["uid"])) == len(d["uid"])           # uid-keyed streams never collide here
    n_before = len(el)
    for q in mp:
        P.droplow(q)
    n_after = len(el)
    assert n_after <= n_before
    assert P.nactives(el) == n_after                # repack!: every surviving row is active
    assert P.weight(el) >= 0.99 * w0
    # idempotence of repack!
    a = el.download(("uid",))["uid"]
    P.repack(el)
    assert np.array_equal(a, el.download(("uid",))["uid"])
    assert gctx.error_flags() == 0


# ---------------------------------------------------------------------------------------------------
# population control: roulette! / split!, and the run! loop
# ---------------------------------------------------------------------------------------------------
def test_roulette_and_split_replay(gctx, octx, air_tables):
    st = _random_pop(np.random.default_rng(12), 50000, 0.1)
    out = []
    for ctx in (gctx, octx):
        ctx.set_rng(21, 5)
        pop = P.Population(ctx, P.ELECTRON, 200000, st, air_tables["electron"], 1e3 * co.eV)
        P.roulette(0.4, pop)
        a = pop.download()
        P.repack(pop)
        P.split(1.5, pop)
        n_split = len(pop)
        b = pop.download()
        P.repack(pop)
        out.append((a, n_split, b, P.weight(pop), len(pop)))
    (ag, ng, bg, wg, lg), (ao, no, bo, wo, lo) = out
    for k in ag:
        assert np.array_equal(ag[k], ao[k]), k                 # roulette!: same survivors, same weights (w / p)
    kept = ag["active"].sum() / st["active"].sum()
    assert abs(kept - 0.4) < 0.02
    assert ng == no and lg == lo
    assert wg == pytest.approx(wo, rel=1e-12)
    # split!: copies carry w / (1 + p); total weight is conserved in expectation
    assert sorted(bg["uid"].tolist()) == sorted(bo["uid"].tolist())
    # (copies of particles at or below the energy cut are refused by add_particle!, population.jl:105)
    alive = ag["active"] == 1
    above = P.kinenergy(P.ELECTRON, ag["p"]) > 1e3 * co.eV
    w_expect = (ag["w"][alive] * (1 / 2.5 + above[alive] * 1.5 / 2.5)).sum()
    assert abs(wg / w_expect - 1) < 0.02


def test_run_loop_observables_agree(gctx, octx, air_tables):
    """run! for 12 steps (advance! + droplow! each step): counts, mean energy, centroid and the energy spectrum of the
    CUDA path and of the oracle agree (same uid-keyed streams => almost particle-identical histories)."""
    res = []
    for ctx in (gctx, octx):
        ctx.set_rng(33, 0)
        mp, el, ph, po = make_world(ctx, air_tables, 4000, 0, 0, cap=60000, seed=13, emin=5e5, emax=2e7)
        P.run(mp, default_pusher(), 12 * DT, DT, P.VoidCallback(), output_dt=None, verbosity=0)
        d = el.download()
        res.append((len(el), len(ph), len(po), P.meanenergy(el), P.spread(el), P.kinenergy(P.ELECTRON, d["p"])))
    (neg, npg, nposg, meg, spg, eg), (neo, npo_, nposo, meo, spo, eo) = res
    assert abs(neg - neo) <= 0.01 * neo + 5
    assert abs(npg - npo_) <= 0.05 * npo_ + 5
    assert meg == pytest.approx(meo, rel=5e-3)
    np.testing.assert_allclose(spg[0], spo[0], rtol=5e-3, atol=1e-4)
    from scipy import stats
    assert stats.ks_2samp(eg, eo).pvalue > 0.01


def test_statistical_agreement_over_seeds(gctx, octx, air_tables):
    """Statistical tier: independent seeds on each side — avalanche multiplication, mean energy and drift of a
    1 MeV electron swarm after 8 steps agree within 3 sigma of the seed-to-seed scatter."""
    def observe(ctx, seed):
        ctx.set_rng(seed, 0)
        mp, el, ph, po = make_world(ctx, air_tables, 1500, 0, 0, cap=30000, seed=seed, espec=1e6)
        P.run(mp, default_pusher(), 8 * DT, DT, P.VoidCallback(), output_dt=None, verbosity=0)
        d = el.download()
        e = P.kinenergy(P.ELECTRON, d["p"])
        hi = e > 1e5 * co.eV
        return [len(el) / 1500.0, float(e[hi].mean() / co.eV), float(d["x"][hi, 2].mean())]
    g = np.array([observe(gctx, 100 + s) for s in range(6)])
    o = np.array([observe(octx, 200 + s) for s in range(6)])
    for k in range(3):
        sigma = np.sqrt(g[:, k].var(ddof=1) / 6 + o[:, k].var(ddof=1) / 6)
        assert abs(g[:, k].mean() - o[:, k].mean()) <= 3 * sigma + 1e-12, (k, g[:, k].mean(), o[:, k].mean(), sigma)


# ---------------------------------------------------------------------------------------------------
# the general forcing / pusher stack (pusher.jl:8-76, field.jl:4-70, continuum.jl) and the remaining processes
# ---------------------------------------------------------------------------------------------------
def _pushers():
    nel = co.nair * 14.4
    cl = P.ContinuumLoss(nel, 85.7 * co.eV, 1e3 * co.eV)
    ccl = P.ChebContinuumLoss.from_loss(cl, 3e8 * co.eV, 4)
    E = [0.0, 0.0, -5e5]
    return {
        "double_layer+B": P.RK2Pusher(P.ElectromagneticField(P.DoubleLayerField(-0.5, 0.5, E), P.HomogeneousField([0.0, 2e-4, 1e-4]))),
        "step": P.RK2Pusher(P.ElectromagneticField(P.StepField(0.0, [1e5, 0.0, -5e5], [0.0, -2e5, 5e5]), None)),
        "confined": P.RK2Pusher(P.ElectromagneticField(P.ConfinedDoubleLayerField(2.0, 3.0, 1.5, -8e5), P.HomogeneousField([0.0, 0.0, 5e-5]))),
        "em+continuum": P.RK2Pusher(P.CombinedForcing(P.ElectromagneticField(P.HomogeneousField(E), P.HomogeneousField([0, 0, 0])), cl)),
        "em+cheb_continuum": P.RK2Pusher(P.CombinedForcing(P.ElectromagneticField(P.HomogeneousField(E), None), ccl)),
        "restricted_forcing": P.RK2Pusher(P.CombinedForcing(P.RestrictedForcing(P.POSITRON, P.ElectromagneticField(P.HomogeneousField(E), None)),
                                                            P.RestrictedForcing(P.ELECTRON, cl))),
        "restricted_pusher": P.RestrictedPusher(P.ELECTRON, P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField(E), None))),
        "null_pusher": P.NullPusher(),
    }


@pytest.mark.parametrize("name", ["double_layer+B", "step", "confined", "em+continuum", "em+cheb_continuum", "restricted_forcing",
                                  "restricted_pusher", "null_pusher"])
def test_forcing_and_pusher_stack_replay(gctx, octx, air_tables, name):
    worlds = []
    for ctx in (gctx, octx):
        ctx.set_rng(17, 0)
        worlds.append(make_world(ctx, air_tables, 1500, 600, 700, cap=30000, seed=21, emin=3e3, emax=2e7))
    psh = _pushers()[name]
    for mp, *_ in worlds:
        P.advance(mp, psh, DT)
    sg, so = P.last_advance_stats(worlds[0][0]), P.last_advance_stats(worlds[1][0])
    assert abs(sg["substeps"] - so["substeps"]) <= 2e-3 * so["substeps"] + 50
    for k, label in ((1, "electron"), (2, "photon"), (3, "positron")):
        _compare_populations(worlds[0][k], worlds[1][k], f"{name}/{label}")
    assert gctx.error_flags(clear=True) == octx.error_flags(clear=True)


def test_moller_and_klein_nishina_tables_replay(gctx, octx):
    """Processes that the air tables of scripts/beam.jl do not use: Moller (moller.jl) with continuum losses below the
    cut, and the closed-form Klein-Nishina cross-section (compton.jl:34-47)."""
    from particulator_b200 import tables
    n2 = co.nair
    et = tables.collision_table_from_processes([(2 * n2, P.RelativisticCoulomb(7)), (14 * n2, P.Moller(1, 1e3 * co.eV))], P.ELECTRON,
                                               co.elementary_charge * 5e5 * DT, safety=1.15)
    gt = tables.collision_table_from_processes([(2 * n2, P.KleinNishinaCompton(7)), (2 * n2, P.PhotoElectric(7))], P.PHOTON, 0)
    for tab, species, lo in ((et, P.ELECTRON, 3e3), (gt, P.PHOTON, 2e3)):
        gctx.set_rng(5, 2)
        octx.set_rng(5, 2)
        for j, proc in enumerate(tab.proc):
            p3 = _momenta(species, 10000, lo, 5e7, 300 + j)
            g = gctx.collide_test(species, tab, j, p3, uid0=77)
            o = octx.collide_test(species, tab, j, p3, uid0=77)
            same = (g[:, 3] == o[:, 3]) & (g[:, 0] == o[:, 0])
            assert (~same).sum() <= 2, proc.name
            scale = np.linalg.norm(p3, axis=1)[same, None]
            for c0 in (4, 8):
                assert (np.abs(g[same, c0:c0 + 3] - o[same, c0:c0 + 3]) / scale).max() <= EVENT_RTOL, proc.name
    # a full step with the Moller table + continuum friction
    cl = P.ContinuumLoss(co.nair * 14.4, 85.7 * co.eV, 1e3 * co.eV)
    psh = P.RK2Pusher(P.CombinedForcing(P.ElectromagneticField(P.HomogeneousField([0, 0, -5e5]), None), cl))
    pops = []
    for ctx in (gctx, octx):
        ctx.set_rng(8, 0)
        rng = np.random.default_rng(31)
        p3 = _momenta(P.ELECTRON, 3000, 5e3, 2e7, 9)
        st = dict(x=np.zeros((3000, 3)), p=p3, s=-np.log(1 - rng.random(3000)), uid=np.arange(1, 3001, dtype=np.uint64))
        el = P.Population(ctx, P.ELECTRON, 30000, st, et, 1e3 * co.eV)
        mp = P.MultiPopulation(("electron", el))
        P.advance(mp, psh, DT)
        pops.append(el)
    _compare_populations(pops[0], pops[1], "moller step")


def test_empty_and_inactive_populations(gctx, air_tables):
    mp, el, ph, po = make_world(gctx, air_tables, 0, 0, 0, cap=1024)
    assert P.advance(mp, default_pusher(), DT) == 0
    assert P.last_advance_stats(mp)["substeps"] == 0
    assert [len(q) for q in mp] == [0, 0, 0]
    assert P.droplow(el) == 0 and P.repack(el) == 0
    assert P.nactives(el) == 0
    # a population whose rows are all inactive is skipped (l.active || continue) and compacts to nothing
    st = _random_pop(np.random.default_rng(0), 500, 1.0)
    el2 = P.Population(gctx, P.ELECTRON, 1000, st, air_tables["electron"], 1e3 * co.eV)
    mp2 = P.MultiPopulation(("electron", el2))
    P.advance(mp2, default_pusher(), DT)
    assert P.last_advance_stats(mp2)["substeps"] == 0
    after = el2.download()
    assert np.array_equal(after["t"], st["t"]) and np.array_equal(after["p"], st["p"])
    assert P.droplow(el2) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [3, 4, 5])
def test_lepton_kernel_variants_are_bit_identical(gctx, air_tables, variant):
    """The three lepton kernels (bq list-scheduled, wf re-sorting, wq warp-private pools) only differ in who executes which
    work unit when: the arithmetic, the draw order and the per-particle Philox streams are shared, so the end state of every
    particle (matched by uid) is the SAME BITS whichever kernel ran.  Reference path: src/mixed_population.jl:56-93."""
    def run(kernel):
        ctx = P.Context(device=0)
        try:
            ctx.set_option("kernel", kernel)
            ctx.set_option("small_pass_rows", 0)      # every pass on the wavefront kernel under test
            ctx.set_rng(17, 0)
            mp, el, ph, po = make_world(ctx, air_tables, 6000, 0, 800, cap=60000, seed=31)
            P.advance(mp, default_pusher(), 2.5e-11)
            out = {}
            for nm, q in (("e", el), ("g", ph), ("p", po)):
                d = q.download()
                o = np.argsort(d["uid"], kind="stable")
                out[nm] = {k: v[o] for k, v in d.items()}
            return out, P.last_advance_stats(mp)["substeps"]
        finally:
            ctx.close()
    ref, sub_ref = run(0)
    got, sub = run(variant)
    assert sub == sub_ref
    for nm in ref:
        assert np.array_equal(ref[nm]["uid"], got[nm]["uid"]), nm
        for k in ("x", "p", "t", "s", "r", "w", "active"):
            assert np.array_equal(ref[nm][k].view(np.uint8), got[nm][k].view(np.uint8)), (nm, k)


# ---------------------------------------------------------------------------------------------------
# API completeness (round 2): energy-dependent roulette!/split!, shuffle!, vector rate bound
# ---------------------------------------------------------------------------------------------------
def _retain_law(eng):
    """Retain probability falling with energy: keep every MeV electron, 20 % of the keV ones."""
    return float(np.clip(0.2 + 0.8 * (np.log10(eng / co.eV) - 3.0) / 3.0, 0.2, 1.0))


@pytest.mark.gpu
def test_energy_dependent_roulette_and_split_replay(gctx, octx, air_tables):
    """roulette!(f, popl) / split!(f, popl) with f a function of the energy (population.jl:291-309, 316-335): the law is
    tabulated by the host; the CUDA path and the oracle make the same per-particle decisions (same uid-keyed draws)."""
    st = _random_pop(np.random.default_rng(5), 40000, 0.05)
    out = []
    for ctx in (gctx, octx):
        ctx.set_rng(3, 9)
        pop = P.Population(ctx, P.ELECTRON, 200000, st, air_tables["electron"], 1e3 * co.eV)
        P.roulette(_retain_law, pop, lo=1e3 * co.eV, hi=1e7 * co.eV, nodes=257)
        a = pop.download()
        P.repack(pop)
        P.split(lambda e: 2.0 if e > 1e6 * co.eV else 0.25, pop, lo=1e3 * co.eV, hi=1e7 * co.eV, nodes=513)
        b = pop.download()
        out.append((a, b))
    (ag, bg), (ao, bo) = out
    assert np.array_equal(ag["active"], ao["active"])
    np.testing.assert_allclose(ag["w"], ao["w"], rtol=1e-12)
    eng = P.kinenergy(P.ELECTRON, st["p"])
    alive0 = st["active"] == 1
    for lo_e, hi_e, expect in [(1e3, 3e3, 0.26), (1e6, 1e7, 1.0)]:
        m = alive0 & (eng > lo_e * co.eV) & (eng < hi_e * co.eV)
        kept = ag["active"][m].mean()
        assert abs(kept - expect) < 0.05, (lo_e, kept)
    assert sorted(bg["uid"].tolist()) == sorted(bo["uid"].tolist())
    assert len(bg["uid"]) > int(ag["active"].sum())


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 1000, 100003])
def test_shuffle_same_permutation(gctx, octx, air_tables, n):
    """shuffle!(popl) (population.jl:266-271): the CUDA path and the oracle apply the SAME permutation (keys from the
    counter-based RNG), every column moves with its row, and the permutation changes from call to call."""
    st = _random_pop(np.random.default_rng(n), n, 0.2)
    res = []
    for ctx in (gctx, octx):
        ctx.set_rng(99, 4)
        pop = P.Population(ctx, P.ELECTRON, n + 8, st, air_tables["electron"], 1e3 * co.eV)
        P.shuffle(pop)
        first = pop.download()
        P.shuffle(pop)
        res.append((first, pop.download()))
    (g1, g2), (o1, o2) = res
    for k in g1:
        assert np.array_equal(g1[k], o1[k]), k
        assert np.array_equal(g2[k], o2[k]), k
    assert sorted(g1["uid"].tolist()) == sorted(st["uid"].tolist())
    src = {u: i for i, u in enumerate(st["uid"].tolist())}
    perm = np.array([src[u] for u in g1["uid"].tolist()])
    assert np.array_equal(g1["p"], st["p"][perm]) and np.array_equal(g1["active"], st["active"][perm])
    if n > 100:
        assert (perm != np.arange(n)).mean() > 0.9 and not np.array_equal(g1["uid"], g2["uid"])


def _vb_table():
    """Linear (LXCat-style) table whose rate bound is a VECTOR on the energy grid (collision_table.jl:35-43)."""
    lin = P.synthetic_lxcat_table(grid_kind=0, nE=512)
    tot = lin.rate.sum(axis=0)                       # includes the explicit null row: constant = maxrate
    real = tot - lin.rate[[i for i, p in enumerate(lin.proc) if type(p).__name__ == "NullCollision"][0]]
    # a bound that hugs the real total rate (x1.2, running maximum over neighbours) instead of the global maximum
    rb = 1.2 * np.maximum.reduce([np.roll(real, k) for k in (-2, -1, 0, 1, 2)])
    rb[:2] = rb[2]; rb[-2:] = rb[-3]
    procs = [p for p in lin.proc if type(p).__name__ != "NullCollision"]
    rate = np.ascontiguousarray(lin.rate[[i for i, p in enumerate(lin.proc) if type(p).__name__ != "NullCollision"]])
    return P.CollisionTable(proc=procs, grid_kind=0, L1=lin.L1, L2=lin.L2, nE=lin.nE, rate=rate, maxrate=float(rb.max()), ratebound=rb)


@pytest.mark.gpu
def test_vector_rate_bound_lookup_and_replay(gctx, octx):
    tab = _vb_table()
    e = np.random.default_rng(8).uniform(0, 99.9, 50000) * co.eV
    rg, bg = gctx.table_eval(tab, e)
    ro, bo = octx.table_eval(tab, e)
    assert np.array_equal(rg.view(np.uint64), ro.view(np.uint64))
    assert np.array_equal(bg.view(np.uint64), bo.view(np.uint64))                    # bit-exact tier
    grid = np.linspace(tab.L1, tab.L2, tab.nE)
    np.testing.assert_allclose(bo, np.interp(e, grid, tab.ratebound), rtol=1e-12)   # the oracle against numpy
    assert (bo >= rg.sum(axis=0) * (1 - 1e-12)).all()
    n = 3000
    rng = np.random.default_rng(4)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * 2 * co.eV / co.electron_mass), s=-np.log(1 - rng.random(n)),
              uid=np.arange(1, n + 1, dtype=np.uint64))
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -100 * co.Td * co.nair]), None))
    out = []
    for ctx in (gctx, octx):
        ctx.set_rng(6, 0)
        pop = P.Population(ctx, P.SLOW_ELECTRON, 3 * n, st, tab, 0.0)
        mp = P.MultiPopulation(("slow", pop))
        for k in range(3):
            P.advance(mp, psh, (k + 1) * 1e-12)
        d = pop.download()
        o = np.argsort(d["uid"], kind="stable")
        out.append(({k: v[o] for k, v in d.items()}, P.last_advance_stats(mp)["substeps"], ctx.error_flags()))
    (g, sg, fg), (o, so, fo) = out
    assert fg == 0 and fo == 0
    assert abs(sg - so) <= 2e-3 * so + 5
    common, ig, io = np.intersect1d(g["uid"], o["uid"], return_indices=True)
    assert len(common) >= 0.995 * len(o["uid"])
    scale = np.maximum(np.linalg.norm(o["p"][io], axis=1), 1e-300)[:, None]
    bad = (np.abs(g["p"][ig] - o["p"][io]) / scale).max(axis=1) > 1e-6
    assert bad.sum() <= max(2, 0.002 * len(common)), int(bad.sum())


# ---------------------------------------------------------------------------------------------------
# invariance to the partition: streams are keyed by uid, never by rank
# ---------------------------------------------------------------------------------------------------
def test_results_do_not_depend_on_how_particles_are_sharded(air_tables):
    """The same particles (same uids, same seed) advanced in ONE context and split over TWO contexts end in identical
    per-uid states, bit for bit, births included: nothing in the path depends on the rank or on the neighbours of a
    particle.  (Wavefront kernel forced on every pass so that both runs take the same code path.)"""
    def world(frac):
        ctx = P.Context(device=0)
        ctx.set_option("small_pass_rows", 0)
        ctx.set_rng(5, 0)
        mp, el, ph, po = make_world(ctx, air_tables, 6000, 3000, 600, cap=80000, seed=77)
        if frac is not None:                      # keep one interleaved half of every species
            for q in (el, ph, po):
                d = q.download()
                keep = (np.arange(len(d["uid"])) % 2) == frac
                q.upload({k: v[keep] for k, v in d.items()})
        return ctx, mp, (el, ph, po)

    def run(ctx, mp, pops):
        t = 0.0
        for _ in range(3):
            t += DT
            P.advance(mp, default_pusher(), t)
            for q in pops:
                P.droplow(q)
        out = [_by_uid(q) for q in pops]
        ctx.close()
        return out

    whole = run(*world(None))
    halves = [run(*world(0)), run(*world(1))]
    for k, name in enumerate(("electron", "photon", "positron")):
        merged = {c: np.concatenate([h[k][c] for h in halves]) for c in whole[k]}
        o = np.argsort(merged["uid"], kind="stable")
        assert np.array_equal(merged["uid"][o], whole[k]["uid"]), name
        for c in whole[k]:
            assert np.array_equal(np.ascontiguousarray(merged[c][o]).view(np.uint8), np.ascontiguousarray(whole[k][c]).view(np.uint8)), (name, c)


def test_mixed_population_replay_at_1e5_per_species(gctx, octx, air_tables):
    """BASELINE configs[3] (mixed e-/gamma/e+ feedback population, all nine processes) at 1e5 particles per species, one
    step, every particle compared with the oracle by uid."""
    worlds = []
    for ctx in (gctx, octx):
        ctx.set_rng(4, 0)
        worlds.append(make_world(ctx, air_tables, 100000, 100000, 100000, cap=400000, seed=44, emin=2e3, emax=3e7))
    for mp, *_ in worlds:
        P.advance(mp, default_pusher(), DT)
    sg, so = P.last_advance_stats(worlds[0][0]), P.last_advance_stats(worlds[1][0])
    assert sg["substeps"] == so["substeps"], (sg["substeps"], so["substeps"])
    assert sg["births"] == so["births"]
    for k, label in ((1, "electron"), (2, "photon"), (3, "positron")):
        _compare_populations(worlds[0][k], worlds[1][k], f"mixed1e5/{label}")
    assert gctx.error_flags(clear=True) == octx.error_flags(clear=True) == 0


# ---------------------------------------------------------------------------------------------------
# statistical tier on BASELINE configs[0] (scripts/swarm.jl): independent seeds on both sides
# ---------------------------------------------------------------------------------------------------
def _weighted_ks(xa, wa, xb, wb):
    """Two-sample Kolmogorov-Smirnov distance between weighted samples."""
    xs = np.concatenate([xa, xb])
    o = np.argsort(xs, kind="stable")
    fa = np.concatenate([wa / wa.sum(), np.zeros(len(xb))])[o].cumsum()
    fb = np.concatenate([np.zeros(len(xa)), wb / wb.sum()])[o].cumsum()
    return float(np.abs(fa - fb).max())


def _swarm_observables(ctx, tables, seed, n0=1500, nsteps=80, every=8):
    """scripts/swarm.jl:28-98 re-expressed for the current API: electrons in a uniform field of
    5e5 V/m along -z in STP air, dt = 2.5e-11 s, 2 ns, roulette back to n0 electrons every `every` steps."""
    ctx.set_rng(seed, 0)
    mp, el, ph, po = make_world(ctx, tables, n0, 0, 0, cap=40 * n0, seed=seed, espec=3e6)
    psh = default_pusher()
    t = 0.0
    hist = []
    for k in range(nsteps):
        t += DT
        P.advance(mp, psh, t)
        for q in (el, ph, po):
            P.droplow(q)
        d = el.diag()
        hist.append((t, d.weight, d.wx[2] / d.weight, d.wenergy / d.weight))
        if (k + 1) % every == 0 and len(el) > n0:
            P.roulette(n0 / len(el), el)
            P.repack(el)
    h = np.array(hist)
    late = h[len(h) // 2:]
    growth = np.polyfit(late[:, 0], np.log(late[:, 1]), 1)[0]        # d ln(W)/dt of the WEIGHTED electron number (roulette-invariant)
    drift = np.polyfit(late[:, 0], late[:, 2], 1)[0]                 # d<z>/dt
    d = el.download()
    a = d["active"] == 1
    e = P.kinenergy(P.ELECTRON, d["p"][a])
    cost = d["p"][a][:, 2] / np.linalg.norm(d["p"][a], axis=1)
    return {"growth": growth, "drift": drift, "meanenergy": float(late[:, 3].mean()), "e": e, "cost": cost, "w": d["w"][a],
            "n_photons": len(ph)}


def test_statistical_tier_swarm_over_16_seeds(air_tables):
    """North-star statistics level: avalanche growth rate, mean energy, drift velocity within 3 sigma of the seed-to-seed
    scatter, and KS distances of the energy and cos(theta) spectra no larger than the distances between independent halves of
    ONE code's seeds — 16 independent seeds per side, 2 ns each, with roulette (the CUDA path never sees the oracle's seeds)."""
    from oracle_backend import oracle_context
    nseed = 16
    g, o = [], []
    for s in range(nseed):
        ctx = P.Context(device=0)
        g.append(_swarm_observables(ctx, air_tables, 1000 + s))
        ctx.close()
        ctx = oracle_context()
        o.append(_swarm_observables(ctx, air_tables, 2000 + s))
        ctx.close()
    report = {}
    for key in ("growth", "drift", "meanenergy"):
        a, b = np.array([r[key] for r in g]), np.array([r[key] for r in o])
        sigma = np.sqrt(a.var(ddof=1) / nseed + b.var(ddof=1) / nseed)
        report[key] = (a.mean(), b.mean(), sigma)
        assert abs(a.mean() - b.mean()) <= 3 * sigma, (key, a.mean(), b.mean(), sigma)
    assert 0 < report["drift"][0] < co.c                              # electrons drift against the field (E along -z), slower than light
    pool = lambda rs, key: np.concatenate([r[key] for r in rs])
    for key in ("e", "cost"):
        d_go = _weighted_ks(pool(g, key), pool(g, "w"), pool(o, key), pool(o, "w"))
        d_gg = _weighted_ks(pool(g[::2], key), pool(g[::2], "w"), pool(g[1::2], key), pool(g[1::2], "w"))
        d_oo = _weighted_ks(pool(o[::2], key), pool(o[::2], "w"), pool(o[1::2], key), pool(o[1::2], "w"))
        report["ks_" + key] = (d_go, d_gg, d_oo)
        # independent halves of one code differ by d_gg / d_oo from seed scatter alone (families of an avalanche are
        # correlated, so the textbook critical value does not apply); the two codes must not differ by more than that
        assert d_go <= 1.5 * max(d_gg, d_oo) + 0.01, (key, d_go, d_gg, d_oo)
    print("[statistics]", {k: tuple(float(f"{x:.5g}") for x in v) for k, v in report.items()})
