"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension):
particulator.jl_b200/csrc/libparticulator_b200.so.  nvcc cross-compiles without a GPU.

The library is five translation units — the C ABI (ptl_api.cu) and the advance kernels of each species
(ptl_adv_species.cu with -DPTL_TU_SPECIES=0..3) — compiled in parallel into build/obj/ and linked with the NCCL runtime
the multi-GPU entry points call (libnccl.so.2, found next to torch or on the system)."""
import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libparticulator_b200.so")
OBJDIR = os.path.join(ROOT, "build", "obj")
HEADERS = ["ptl_common.cuh", "ptl_physics.cuh", "ptl_advance.cuh", "ptl_advance_wf.cuh", "ptl_advance_bq.cuh", "ptl_advance_wq.cuh", "ptl_store.cuh",
           "ptl_host.h", "ptl_launch.cuh", os.path.join("..", "..", "include", "particulator_b200.h")]
# (object name, source, extra flags)
UNITS = [("ptl_api", "ptl_api.cu", []), ("ptl_comm", "ptl_comm.cu", [])] + \
        [(f"ptl_adv_sp{k}", "ptl_adv_species.cu", [f"-DPTL_TU_SPECIES={k}"]) for k in range(4)]
SOURCES = sorted({u[1] for u in UNITS})

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f)))


def nccl_paths():
    """Header directory and shared object of the NCCL runtime to link against (torch's bundled copy first)."""
    inc, lib = None, None
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            base = list(spec.submodule_search_locations)[0]
            if os.path.exists(os.path.join(base, "include", "nccl.h")):
                inc = os.path.join(base, "include")
            cands = glob.glob(os.path.join(base, "lib", "libnccl.so*"))
            if cands:
                lib = sorted(cands)[0]
    except Exception:
        pass
    for d in ("/usr/include", "/usr/local/cuda/include"):
        if inc is None and os.path.exists(os.path.join(d, "nccl.h")):
            inc = d
    if lib is None:
        for pat in ("/usr/lib/x86_64-linux-gnu/libnccl.so*", "/usr/local/cuda/lib64/libnccl.so*"):
            cands = glob.glob(pat)
            if cands:
                lib = sorted(cands)[0]
                break
    return inc, lib


def build(force=False, verbose=False, extra=(), lib=LIB, jobs=None):
    if not force and not needs_build(lib):
        return lib
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a gcc wrapper that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    tag = hashlib.sha1((" ".join(extra) + lib).encode()).hexdigest()[:10]
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    inc, _ = nccl_paths()
    incflags = ["-I", inc] if inc else []

    def compile_one(unit):
        name, src, flags = unit
        obj = os.path.join(objdir, name + ".o")
        cmd = [nvcc] + NVCC_FLAGS + incflags + list(extra) + flags + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, env=env)
        return name, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or min(len(UNITS), os.cpu_count() or 1)) as ex:
        for name, obj, r in ex.map(compile_one, UNITS):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"---- {name}\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed compiling {name}")
            objs.append(obj)
    # libnccl is NOT linked: ptl_comm.cu binds it at run time with dlopen (torch has usually loaded it already), so the
    # library also loads on a box without NCCL and the single-GPU path never touches it
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", lib] + objs + ["-ldl"]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libparticulator_b200.so")
    return lib


if __name__ == "__main__":
    args = sys.argv[1:]
    out = LIB
    if "-o" in args:
        k = args.index("-o")
        out = os.path.abspath(args[k + 1])
        del args[k:k + 2]
    print(build(force=True, verbose="--verbose" in args, extra=[a for a in args if a not in ("--force", "--verbose")], lib=out))
