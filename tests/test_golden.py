"""Golden regression vectors (tests/golden/oracle_vectors.npz, written by tests/golden/make_golden.py from the CPU
oracle).  The CPU suite checks that the oracle still reproduces them bit for bit; the GPU suite checks the CUDA path
against the same file (bit-exact lookups, 1e-9 single events, 1e-6 per-particle end states matched by uid)."""
import os
import sys

import numpy as np
import pytest

import particulator_b200 as P

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = dict(np.load(os.path.join(HERE, "golden", "oracle_vectors.npz")))


def test_oracle_reproduces_golden_vectors(octx):
    res = make_golden.compute(octx)
    assert set(res) == set(GOLD)
    for k, v in res.items():
        assert np.array_equal(np.asarray(v), GOLD[k]), k


@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors(gctx):
    res = make_golden.compute(gctx)
    assert set(res) == set(GOLD)
    for k, g in GOLD.items():
        v = np.asarray(res[k])
        if k.startswith("lookup_"):
            assert np.array_equal(v.view(np.uint64), g.view(np.uint64)), k          # bit-exact tier
        elif k.startswith("collide_"):
            assert np.array_equal(v[:, :4], g[:, :4]), k                              # outcome kinds and draw counts
            scale = np.maximum(np.abs(g[:, 4:]).max(axis=1, keepdims=True), 1e-300)
            assert (np.abs(v[:, 4:] - g[:, 4:]) / scale).max() <= 1e-9, k
    assert int(res["advance_substeps"][0]) == int(GOLD["advance_substeps"][0])
    for nm in ("electron", "photon", "positron"):
        ug, uo = res[f"advance_{nm}_uid"], GOLD[f"advance_{nm}_uid"]
        assert np.array_equal(ug, uo), nm
        assert np.array_equal(res[f"advance_{nm}_active"], GOLD[f"advance_{nm}_active"])
        pg, po = res[f"advance_{nm}_p"], GOLD[f"advance_{nm}_p"]
        scale = np.maximum(np.linalg.norm(po, axis=1, keepdims=True), 1e-300)
        assert (np.abs(pg - po) / scale).max() <= 1e-6, nm
        np.testing.assert_allclose(res[f"advance_{nm}_x"], GOLD[f"advance_{nm}_x"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(res[f"advance_{nm}_s"], GOLD[f"advance_{nm}_s"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(res[f"advance_{nm}_r"], GOLD[f"advance_{nm}_r"], rtol=1e-6, atol=1e-3)
