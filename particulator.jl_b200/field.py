"""Analytic fields and the Lorentz forcing — descriptors mirroring src/field.jl.

HomogeneousField :4-8, DoubleLayerField :10-16, StepField :22-28, ConfinedDoubleLayerField :35-52,
ElectromagneticField :54-70.  The evaluation happens on the device (csrc/physics.cuh)."""
from dataclasses import dataclass
from typing import Sequence

from ._lib import FieldDesc, ForcingDesc


def _fd(kind, par):
    d = FieldDesc()
    d.kind = kind
    for i, v in enumerate(par):
        d.par[i] = float(v)
    return d


@dataclass
class ZeroField:
    def desc(self):
        return _fd(0, [])


@dataclass
class HomogeneousField:
    v: Sequence[float]

    def desc(self):
        return _fd(1, list(self.v))


@dataclass
class DoubleLayerField:
    z1: float
    z2: float
    v: Sequence[float]

    def desc(self):
        return _fd(2, [self.z1, self.z2, *self.v])


@dataclass
class StepField:
    z: float
    v1: Sequence[float]
    v2: Sequence[float]

    def desc(self):
        return _fd(3, [self.z, *self.v1, *self.v2])


@dataclass
class ConfinedDoubleLayerField:
    sx: float
    sy: float
    sz: float
    ez0: float

    def desc(self):
        return _fd(4, [self.sx, self.sy, self.sz, self.ez0])


@dataclass
class ElectromagneticField:
    """field.jl:54-57: force = charge * e * (E(x,t) + v x B(x,t)) on e-/e+, zero on photons."""
    e: object
    b: object = None

    def forcing_desc(self, ctx, mask=0):
        f = ForcingDesc()
        f.kind = 1
        f.species_mask = mask
        f.e = self.e.desc()
        f.b = (self.b or ZeroField()).desc()
        return f
