"""The C-ABI shared library loads and exports every symbol include/particulator_b200.h declares
(no compute calls here: this runs without a GPU)."""
import ctypes
import os
import re

import particulator_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "particulator_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ptl_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    assert "ptl_advance" in syms and "ptl_droplow" in syms and "ptl_population_upload" in syms
    assert len(syms) >= 38


def test_cuda_library_exports_every_declared_symbol():
    assert os.path.exists(P.LIB_PATH), "CUDA library not built: run __graft_entry__.build()"
    dll = ctypes.CDLL(P.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(dll, s)]
    assert not missing, f"missing exports: {missing}"
    dll.ptl_abi_version.restype = ctypes.c_int32
    assert dll.ptl_abi_version() == 1


def test_python_binding_covers_the_header():
    declared = {s[len("ptl_"):] for s in _declared_symbols()}
    assert declared == set(P.ABI_SYMBOLS)


def test_oracle_exports_the_same_abi():
    from oracle_backend import oracle_backend
    b = oracle_backend()
    for s in P.ABI_SYMBOLS_CORE:
        assert hasattr(b.dll, "ora_" + s)


def test_no_cpu_fallback_in_product_path():
    """Without a GPU the product context must fail loudly (PTL_ENODEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        P.Context(device=0)
    except P.PtlError as e:
        assert "no CPU fallback" in str(e) or "status" in str(e)
    else:
        raise AssertionError("Context creation must fail without a GPU")


def test_package_does_not_reference_oracle():
    """The product package never imports, links, includes or loads anything under oracle/."""
    pkg = os.path.join(ROOT, "particulator.jl_b200")
    bad = re.compile(r"libptl_oracle|ptl_oracle\.|oracle_backend|import\s+oracle|from\s+oracle|#include\s+\"[^\"]*oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f


def test_bench_shard_bounds_partition_the_rows():
    """bench.py's e2e leg cuts the host arrays into shards: every row in exactly one shard, small shards at both ends."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n, ns, nw, ramp in ((100_000_007, 12, 4, 0.3), (1000, 12, 3, 0.3), (12345, 10, 4, 0.3), (7, 12, 4, 0.3), (5_000_000, 16, 4, 1.0)):
        b = bench.shard_bounds(n, ns, nw, ramp)
        assert len(b) == ns + 1 and b[0] == 0 and b[-1] == n
        assert all(b[k] <= b[k + 1] for k in range(ns))
        sizes = [b[k + 1] - b[k] for k in range(ns)]
        assert sum(sizes) == n
        if ns >= 3 * nw and ramp < 1 and n > 100 * ns:
            mid = sizes[ns // 2]
            assert sizes[0] < 0.5 * mid and sizes[-1] < 0.5 * mid
            assert all(abs(sizes[k] - mid) <= 2 for k in range(nw, ns - nw))
