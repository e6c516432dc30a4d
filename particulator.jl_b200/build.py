"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension):
particulator.jl_b200/csrc/libparticulator_b200.so.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libparticulator_b200.so")
SOURCES = ["ptl_api.cu"]
HEADERS = ["ptl_common.cuh", "ptl_physics.cuh", "ptl_advance.cuh", "ptl_advance_wf.cuh", "ptl_advance_aq.cuh", "ptl_advance_bq.cuh", "ptl_store.cuh",
           os.path.join("..", "..", "include", "particulator_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-o", LIB] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a gcc wrapper that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libparticulator_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, extra=[a for a in sys.argv[1:] if a != "--force"]))
