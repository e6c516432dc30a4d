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


def _beam():
    spec = importlib.util.spec_from_file_location("example_beam", os.path.join(ROOT, "examples", "beam.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


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
