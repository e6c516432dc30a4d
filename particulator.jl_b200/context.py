"""Device context + table upload: the host object that owns one GPU-side library context.

There is no counterpart in the reference (a Julia process simply owns its arrays); the context is
the handle every C-ABI call takes first (include/particulator_b200.h, section "context")."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PtlError, ProcessDesc, dptr, as_f64
from .tables import ChebyshevCollisionTable, CollisionTable, ChebContinuumLoss
from .processes import SeltzerBerger


class Context:
    def __init__(self, device=0, stream=None, backend=None):
        self.backend = backend or _lib.cuda_backend()
        h = C.c_void_p()
        rc = self.backend.context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise PtlError(f"ptl_context_create failed with status {rc} "
                           f"({'no sm_100 CUDA device: there is no CPU fallback' if rc == -2 else 'see status codes'})")
        self.h = h
        self.device = device
        self._tables = {}
        self._sb = {}
        self._cheb_loss = {}

    def close(self):
        if self.h:
            self.backend.context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- status ---------------------------------------------------------------------------------
    def check(self, rc, what=""):
        if rc < 0:
            msg = self.backend.last_error(self.h)
            raise PtlError(f"{what}: usage error {rc}: {msg.decode() if msg else ''}")
        return rc

    def raise_on_flags(self, rc, what=""):
        """Turn sticky device conditions into the exceptions the reference would throw
        (AssertionError at population.jl:107, collisions.jl:186, ...)."""
        self.check(rc, what)
        if rc > 0:
            raise PtlError(f"{what}: device condition {_lib.describe_flags(rc)}")

    def error_flags(self, clear=False):
        return self.backend.error_flags(self.h, 1 if clear else 0)

    def set_profiling(self, on=True):
        self.check(self.backend.set_profiling(self.h, 1 if on else 0), "set_profiling")

    def launch_count(self, reset=False):
        return int(self.backend.launch_count(self.h, 1 if reset else 0))

    def synchronize(self):
        self.check(self.backend.synchronize(self.h), "synchronize")

    def set_rng(self, seed, step=0):
        self.check(self.backend.set_rng(self.h, int(seed), int(step)), "set_rng")

    def get_rng(self):
        s, st = C.c_uint64(), C.c_uint32()
        self.backend.get_rng(self.h, C.byref(s), C.byref(st))
        return s.value, st.value

    def get_uid_counter(self):
        return int(self.backend.get_uid_counter(self.h))

    def set_uid_counter(self, next_uid):
        self.check(self.backend.set_uid_counter(self.h, int(next_uid)), "set_uid_counter")

    def set_option(self, name, value):
        """Tuning knobs outside the reference's surface ("kernel": lepton kernel variant, "stream": streaming fast path)."""
        self.check(self.backend.set_option(self.h, str(name).encode(), int(value)), "set_option")

    # -- tables ---------------------------------------------------------------------------------
    def _proc_descs(self, procs):
        arr = (ProcessDesc * max(1, len(procs)))()
        for i, p in enumerate(procs):
            arr[i].kind = p.kind
            if isinstance(p, SeltzerBerger):
                arr[i].aux = self.sb_table(p)
            else:
                arr[i].aux = -1
            par = p.params()
            for k in range(_lib.PROC_NPAR):
                arr[i].par[k] = par[k] if k < len(par) else 0.0
        return arr

    def sb_table(self, sb: SeltzerBerger):
        key = id(sb)
        if key not in self._sb:
            data = np.asfortranarray(sb.data)
            le = as_f64(sb.log_energy)
            rc = self.backend.sb_table_create(self.h, data.shape[0], data.shape[1], dptr(le),
                                              data.ctypes.data_as(C.POINTER(C.c_double)))
            self._sb[key] = (self.check(rc, "sb_table_create"), sb)
        return self._sb[key][0]

    def table(self, tab):
        """Upload a host-built collision table once; returns its id."""
        key = id(tab)
        if key in self._tables:
            return self._tables[key][0]
        procs = self._proc_descs(tab.proc)
        if isinstance(tab, ChebyshevCollisionTable):
            rate = np.asfortranarray(tab.rate)          # [order, nprocs, k+1], order fastest
            rb = np.asfortranarray(tab.ratebound)
            rc = self.backend.table_create_cheb(self.h, tab.order, len(tab.proc), tab.b.k, float(tab.b.xmax),
                                                rate.ctypes.data_as(C.POINTER(C.c_double)),
                                                rb.ctypes.data_as(C.POINTER(C.c_double)), procs)
        elif isinstance(tab, CollisionTable):
            rate = np.asfortranarray(tab.rate)          # [nprocs, nE], process fastest
            if tab.ratebound is not None:               # vector rate bound (collision_table.jl:35-43)
                rb = as_f64(tab.ratebound).ravel()
                assert len(rb) == tab.nE
                rc = self.backend.table_create_linear_vb(self.h, tab.grid_kind, float(tab.L1), float(tab.L2), tab.nE,
                                                         len(tab.proc), rate.ctypes.data_as(C.POINTER(C.c_double)), dptr(rb), procs)
            else:
                rc = self.backend.table_create_linear(self.h, tab.grid_kind, float(tab.L1), float(tab.L2), tab.nE,
                                                      len(tab.proc), rate.ctypes.data_as(C.POINTER(C.c_double)),
                                                      float(tab.maxrate), procs)
        else:
            raise TypeError(type(tab))
        tid = self.check(rc, "table_create")
        self._tables[key] = (tid, tab)
        return tid

    def cheb_loss(self, cl: ChebContinuumLoss):
        key = id(cl)
        if key not in self._cheb_loss:
            ec, pc = np.asfortranarray(cl.ec), np.asfortranarray(cl.pc)
            rc = self.backend.cheb_loss_create(self.h, ec.shape[0], cl.bints.k, float(cl.bints.xmax),
                                               ec.ctypes.data_as(C.POINTER(C.c_double)),
                                               pc.ctypes.data_as(C.POINTER(C.c_double)))
            self._cheb_loss[key] = (self.check(rc, "cheb_loss_create"), cl)
        return self._cheb_loss[key][0]

    def table_eval(self, tab, energy):
        """rate(table, j, presample(E)) for every process j and ratebound(E), evaluated by the library
        (collision_table.jl:50-57,82-106).  Returns (rates[nprocs, n], bound[n])."""
        tid = self.table(tab)
        e = as_f64(energy).ravel()
        rates = np.zeros((len(tab.proc), len(e)), order="F")
        bound = np.zeros(len(e))
        rc = self.backend.table_eval(self.h, tid, len(e), dptr(e), rates.ctypes.data_as(C.POINTER(C.c_double)), dptr(bound))
        self.check(rc, "table_eval")
        return rates, bound

    def rng_test(self, uid, seed, step, n):
        out = np.zeros(n)
        self.check(self.backend.rng_test(self.h, int(uid), int(seed), int(step), n, dptr(out)), "rng_test")
        return out

    def collide_test(self, species, tab, j, p3, uid0=1):
        tid = self.table(tab)
        p3 = as_f64(p3).reshape(-1, 3)
        out = np.zeros((p3.shape[0], 24))
        rc = self.backend.collide_test(self.h, species, tid, j, p3.shape[0], dptr(p3), int(uid0), dptr(out))
        self.check(rc, "collide_test")
        return out
