import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import particulator_b200 as P  # noqa: E402
from oracle_backend import oracle_context  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def air_tables():
    """Electron / positron / photon tables of scripts/beam.jl:94-129 for E = 5e5 V/m, dt = 2.5e-11 s."""
    co = P.co
    comp = P.air_composition()
    Fdt = co.elementary_charge * 5e5 * 2.5e-11
    return {
        "electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
        "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
        "photon": P.build_photon_collision_table(comp),
    }


@pytest.fixture()
def octx():
    ctx = oracle_context()
    yield ctx
    ctx.close()


@pytest.fixture(params=["wavefront", "default"])
def gctx(request):
    """CUDA context.  No fallback: if the library or the device is missing the test FAILS.
    Every GPU test runs twice: with the wavefront lepton kernel forced for every pass ("wavefront": small_pass_rows = 0, so
    that test-sized populations exercise the kernel the benchmark runs), and with the library's default dispatch, which sends
    passes of fewer than 16384 rows to the one-particle-per-lane kernel."""
    ctx = P.Context(device=0)
    if request.param == "wavefront":
        ctx.set_option("small_pass_rows", 0)
    yield ctx
    ctx.close()


def make_world(ctx, tables, n_e=0, n_g=0, n_p=0, cap=None, seed=0, cuts=(1e3, 1e3, 1e2), espec="log", emin=2e3, emax=5e7):
    """Three populations (electron, photon, positron) with seeded synthetic particles."""
    co = P.co
    rng = np.random.default_rng(seed)

    def mk(species, n, lo, hi):
        if n == 0:
            return None
        if espec == "log":
            K = np.exp(rng.uniform(np.log(lo), np.log(hi), n)) * co.eV
        else:
            K = np.full(n, espec) * co.eV
        pn = P.momentum_norm_from_kin(species, K)
        cost = rng.uniform(-1, 1, n)
        phi = rng.uniform(0, 2 * np.pi, n)
        sint = np.sqrt(1 - cost ** 2)
        d = np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)
        return dict(x=rng.normal(0, 1.0, (n, 3)), p=d * pn[:, None], s=-np.log(1 - rng.random(n)),
                    uid=np.arange(1, n + 1, dtype=np.uint64) + np.uint64(species * 10 ** 9))

    cap = cap or max(1024, 4 * (n_e + n_g + n_p))
    el = P.Population(ctx, P.ELECTRON, cap, mk(P.ELECTRON, n_e, emin, emax), tables["electron"], cuts[0] * co.eV)
    ph = P.Population(ctx, P.PHOTON, cap, mk(P.PHOTON, n_g, max(emin, 2e3), emax), tables["photon"], cuts[1] * co.eV)
    po = P.Population(ctx, P.POSITRON, cap, mk(P.POSITRON, n_p, emin, emax), tables["positron"], cuts[2] * co.eV)
    mp = P.MultiPopulation(("electron", el), ("photon", ph), ("positron", po))
    return mp, el, ph, po


def default_pusher():
    return P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -5e5]), P.HomogeneousField([0.0, 0.0, 0.0])))
