"""The example script that mirrors the reference's scripts/beam.jl runs end to end (oracle backend on CPU, CUDA on the GPU box)
and its observables agree between the two."""
import importlib.util
import os

import numpy as np
import pytest

import particulator_b200 as P
from oracle_backend import oracle_context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
co = P.co


def _load(name):
    spec = importlib.util.spec_from_file_location("example_" + name, os.path.join(ROOT, "examples", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _beam():
    return _load("beam")


def _summary(r):
    el, ph = r["electrons"], r["photons"]
    return len(el), len(ph), P.meanenergy(el), P.spread(el)[0][2]


def test_beam_example_runs_on_the_oracle():
    r = _beam().main(n_init_particles=20, maxp=20000, tfinal=2.5e-10, zwall=None, ctx=oracle_context())
    n, ng, emean, z = _summary(r)
    assert n >= 20 and 2.5e-10 - 1e-20 <= r["t"] <= 2.5e-10 + 2.5e-11 * 1.0001   # run!: `while t < tfinal` on an accumulated t (run.jl:5)
    assert 0.0 < z < 0.1                        # 0.25 ns at ~c: centroid near 7 cm, secondaries pull it back
    assert emean < 7e6 * co.eV


@pytest.mark.gpu
def test_beam_example_gpu_matches_oracle():
    b = _beam()
    rg = b.main(n_init_particles=200, maxp=100000, tfinal=5e-10, ctx=P.Context(device=0))
    ro = b.main(n_init_particles=200, maxp=100000, tfinal=5e-10, ctx=oracle_context())
    ng, no = _summary(rg), _summary(ro)
    assert abs(ng[0] - no[0]) <= 2 + 0.01 * no[0] and abs(ng[1] - no[1]) <= 2 + 0.02 * no[1]
    assert ng[2] == pytest.approx(no[2], rel=1e-3) and ng[3] == pytest.approx(no[3], rel=1e-3)


def test_swarm_example_roulette_keeps_the_weighted_count():
    """scripts/swarm.jl: roulette! back to ntarget after every outer iteration; the weighted count follows the avalanche."""
    r = _load("swarm").main(n_init_particles=60, maxp=50000, ntarget=40, iterations=3, tstep=1e-10, ctx=oracle_context())
    h = r["history"]
    assert len(h) == 3 and all(n > 40 for _, n, _ in h[:1])
    assert P.nactives(r["electrons"]) <= 60                       # rouletted
    assert h[-1][2] >= 0.5 * h[0][1]                              # weighted count is not lost by the roulette (p = ntarget/n, w /= p)


@pytest.mark.gpu
def test_swarm_example_runs_on_gpu():
    r = _load("swarm").main(n_init_particles=2000, maxp=200000, ntarget=1500, iterations=4, tstep=2e-10, ctx=P.Context(device=0))
    h = r["history"]
    assert len(h) == 4 and P.nactives(r["electrons"]) <= 2200
    assert h[-1][2] > 1500 and np.isfinite(h[-1][2])              # weighted electrons keep growing past the target
