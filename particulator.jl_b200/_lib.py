"""ctypes binding of the C ABI declared in include/particulator_b200.h.

`Backend` binds one shared library that exports the ABI under a symbol prefix.  The product
backend is `cuda_backend()` — libparticulator_b200.so, prefix `ptl_` — and there is no other
backend in this package: if the CUDA library is missing or no sm_100 device is present the
call fails loudly.  (tests/ builds a second `Backend` over the CPU oracle, prefix `ora_`, to
drive both implementations through the same host classes; the package itself never does.)"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTL_LIB_PATH") or os.path.join(_HERE, "csrc", "libparticulator_b200.so")   # env override: A/B builds of the same CUDA library

MAX_PROCS = 128
PROC_NPAR = 6
MAX_FORCINGS = 4
MAX_WALLS = 4


class ProcessDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("aux", C.c_int32), ("par", C.c_double * PROC_NPAR)]


class FieldDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("par", C.c_double * 7)]


class ForcingDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species_mask", C.c_uint32), ("e", FieldDesc), ("b", FieldDesc),
                ("nel", C.c_double), ("I", C.c_double), ("Tcut", C.c_double), ("cheb_id", C.c_int32), ("_pad", C.c_int32)]


class PusherDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("restrict_mask", C.c_uint32), ("nforcings", C.c_int32), ("_pad", C.c_int32),
                ("forcing", ForcingDesc * MAX_FORCINGS)]


class WallDesc(C.Structure):
    _fields_ = [("species", C.c_int32), ("coord", C.c_int32), ("v", C.c_double), ("drop", C.c_int32), ("_pad", C.c_int32)]


class CallbackDesc(C.Structure):
    _fields_ = [("nwalls", C.c_int32), ("count_collisions", C.c_int32), ("wall", WallDesc * MAX_WALLS)]


class DiagOut(C.Structure):
    _fields_ = [("n", C.c_int64), ("nactive", C.c_int64), ("weight", C.c_double), ("wenergy", C.c_double),
                ("maxenergy", C.c_double), ("wx", C.c_double * 3), ("wx2", C.c_double * 3), ("wr2", C.c_double)]


class AdvanceStats(C.Structure):
    _fields_ = [("passes", C.c_int64), ("substeps", C.c_int64), ("rows", C.c_int64), ("births", C.c_int64),
                ("launches", C.c_int64), ("main_rows", C.c_int64), ("main_ms", C.c_double)]


_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); the first argument (context pointer) is implicit unless noted
_SIGS = {
    "abi_version": (C.c_int32, None),
    "context_create": (C.c_int32, [C.c_int32, _vp, C.POINTER(_vp)]),
    "context_destroy": (C.c_int32, [_vp]),
    "last_error": (C.c_char_p, [_vp]),
    "error_flags": (C.c_int32, [_vp, C.c_int32]),
    "synchronize": (C.c_int32, [_vp]),
    "set_rng": (C.c_int32, [_vp, C.c_uint64, C.c_uint32]),
    "get_rng": (C.c_int32, [_vp, _u64p, C.POINTER(C.c_uint32)]),
    "sb_table_create": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp, _dp]),
    "table_create_cheb": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp, _dp, C.POINTER(ProcessDesc)]),
    "table_create_linear": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp, C.c_double,
                                        C.POINTER(ProcessDesc)]),
    "cheb_loss_create": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_double, _dp, _dp]),
    "table_eval": (C.c_int32, [_vp, C.c_int32, C.c_int64, _dp, _dp, _dp]),
    "population_create": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.c_double, C.c_int32]),
    "population_destroy": (C.c_int32, [_vp, C.c_int32]),
    "population_upload": (C.c_int32, [_vp, C.c_int32, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _u8p, _u64p]),
    "population_download": (C.c_int64, [_vp, C.c_int32, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _u8p, _u64p]),
    "population_n": (C.c_int64, [_vp, C.c_int32]),
    "population_capacity": (C.c_int64, [_vp, C.c_int32]),
    "population_clear": (C.c_int32, [_vp, C.c_int32]),
    "population_append": (C.c_int64, [_vp, C.c_int32, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64]),
    "population_deactivate": (C.c_int32, [_vp, C.c_int32, C.c_int64]),
    "droplow": (C.c_int64, [_vp, C.c_int32, C.c_double]),
    "repack": (C.c_int64, [_vp, C.c_int32]),
    "diag": (C.c_int32, [_vp, C.c_int32, C.POINTER(DiagOut)]),
    "histogram": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp]),
    "roulette": (C.c_int32, [_vp, C.c_int32, C.c_double]),
    "split": (C.c_int32, [_vp, C.c_int32, C.c_double]),
    "population_column_ptr": (_vp, [_vp, C.c_int32, C.c_int32]),
    "population_set_n": (C.c_int32, [_vp, C.c_int32, C.c_int64]),
    "multipop_create": (C.c_int32, [_vp, C.POINTER(C.c_int32), C.c_int32]),
    "init": (C.c_int32, [_vp, C.c_int32]),
    "advance": (C.c_int32, [_vp, C.c_int32, C.POINTER(PusherDesc), C.c_double, C.POINTER(CallbackDesc)]),
    "last_advance_stats": (C.c_int32, [_vp, C.POINTER(AdvanceStats)]),
    "set_profiling": (C.c_int32, [_vp, C.c_int32]),
    "launch_count": (C.c_int64, [_vp, C.c_int32]),
    "collision_counts": (C.c_int32, [_vp, C.c_int32, _i64p, C.c_int32]),
    "wall_records": (C.c_int64, [_vp, C.c_int32, C.c_int64, _dp, _dp, _dp, _dp, C.c_int32]),
    "collide_test": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _dp, C.c_uint64, _dp]),
    "rng_test": (C.c_int32, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, _dp]),
    "table_create_linear_vb": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp, _dp,
                                           C.POINTER(ProcessDesc)]),
    "roulette_law": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp]),
    "split_law": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp]),
    "shuffle": (C.c_int32, [_vp, C.c_int32]),
    "set_uid_counter": (C.c_int32, [_vp, C.c_uint64]),
    "get_uid_counter": (C.c_uint64, [_vp]),
}

# entry points only the CUDA library has: the multi-GPU communicator (NCCL) and tuning knobs.  The CPU oracle the tests
# drive through the same host classes is a single-process checker and does not export them.
_SIGS_DEVICE_ONLY = {
    "set_option": (C.c_int32, [_vp, C.c_char_p, C.c_int64]),
    "comm_unique_id": (C.c_int32, [_u8p]),
    "comm_init": (C.c_int32, [_vp, _u8p, C.c_int32, C.c_int32]),
    "comm_destroy": (C.c_int32, [_vp]),
    "comm_info": (C.c_int32, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "diag_allreduce": (C.c_int32, [_vp, C.c_int32, C.POINTER(DiagOut)]),
    "histogram_allreduce": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp]),
    "comm_allreduce_f64": (C.c_int32, [_vp, _dp, C.c_int32, C.c_int32]),
    "rebalance_plan": (C.c_int32, [_i64p, C.c_int32, C.c_double, _i64p, C.c_int32]),
    "rebalance": (C.c_int64, [_vp, C.c_int32, C.c_double, _i64p]),
}

#: every symbol include/particulator_b200.h declares (suffix after the prefix)
ABI_SYMBOLS = sorted(list(_SIGS) + list(_SIGS_DEVICE_ONLY))
#: the subset both the CUDA library and the test-side CPU oracle export
ABI_SYMBOLS_CORE = sorted(_SIGS)


class PtlError(RuntimeError):
    pass


ERR_BITS = {1: "CAPACITY_OVERFLOW", 2: "RATE_BOUND_VIOLATED", 4: "ENERGY_OUT_OF_TABLE", 8: "NAN_STATE",
            16: "SAMPLER_INVARIANT"}


def describe_flags(flags):
    return "|".join(name for bit, name in ERR_BITS.items() if flags & bit) or "0"


class Backend:
    """One shared library exporting the ABI under `prefix`."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise PtlError(f"shared library not found: {path} (build it: python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self.fn = {}
        sigs = dict(_SIGS)
        sigs.update({k: v for k, v in _SIGS_DEVICE_ONLY.items() if hasattr(self.dll, prefix + k)})
        tolerant = bool(os.environ.get("PTL_AB_OLD_LIB"))     # development aid: A/B against a library built from an older commit
        for name, (res, args) in sigs.items():
            if tolerant and not hasattr(self.dll, prefix + name):
                continue
            f = getattr(self.dll, prefix + name)
            f.restype = res
            if args is not None:
                f.argtypes = args
            self.fn[name] = f

    def __getattr__(self, name):
        try:
            return self.__dict__["fn"][name]
        except KeyError:
            raise AttributeError(name)


_cuda = None


def cuda_backend():
    """The product backend.  Raises if the CUDA library has not been built — there is no fallback."""
    global _cuda
    if _cuda is None:
        _cuda = Backend(LIB_PATH, "ptl_")
    return _cuda


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a
