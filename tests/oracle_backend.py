"""Test-side helper: a `Backend` over the CPU oracle (oracle/libptl_oracle.so, prefix ora_).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this."""
import os
import subprocess

import particulator_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libptl_oracle.so")

_backend = None


def build_oracle():
    src = os.path.join(ORACLE_DIR, "ptl_oracle.c")
    if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_LIB


def oracle_backend():
    global _backend
    if _backend is None:
        _backend = P.Backend(build_oracle(), "ora_")
    return _backend


def oracle_context():
    return P.Context(backend=oracle_backend())
