"""Pushers and forcings — descriptors mirroring src/pusher.jl and src/continuum.jl.

NullForcing :11, CombinedForcing :14-23, RestrictedForcing :28-34, RK2Pusher :37-63,
RestrictedPusher :67-73, NullPusher :75-76; ContinuumLoss / ChebContinuumLoss continuum.jl:6-57."""
from dataclasses import dataclass

from ._lib import PusherDesc, ForcingDesc, MAX_FORCINGS
from . import tables


class NullForcing:
    def forcing_desc(self, ctx, mask=0):
        f = ForcingDesc()
        f.kind = 0
        return f


class CombinedForcing:
    def __init__(self, *terms):
        self.tpl = tuple(terms)


class RestrictedForcing:
    """RestrictedForcing{T}(forcing): acts on one species only."""

    def __init__(self, species, forcing):
        self.species, self.forcing = species, forcing

    def forcing_desc(self, ctx, mask=0):
        return self.forcing.forcing_desc(ctx, mask=1 << self.species)


def _continuum_desc(self, ctx, mask=0):
    f = ForcingDesc()
    f.kind = 2
    f.species_mask = mask
    f.nel, f.I, f.Tcut = self.nel, self.I, self.Tcut
    return f


def _cheb_continuum_desc(self, ctx, mask=0):
    f = ForcingDesc()
    f.kind = 3
    f.species_mask = mask
    f.cheb_id = ctx.cheb_loss(self)
    return f


tables.ContinuumLoss.forcing_desc = _continuum_desc
tables.ChebContinuumLoss.forcing_desc = _cheb_continuum_desc
ContinuumLoss = tables.ContinuumLoss
ChebContinuumLoss = tables.ChebContinuumLoss


def _flatten(forcing):
    if forcing is None:
        return []
    if isinstance(forcing, CombinedForcing):
        out = []
        for t in forcing.tpl:
            out += _flatten(t)
        return out
    return [forcing]


@dataclass
class RK2Pusher:
    forcing: object

    def desc(self, ctx, restrict_mask=0):
        d = PusherDesc()
        d.kind = 1
        d.restrict_mask = restrict_mask
        terms = _flatten(self.forcing)
        if len(terms) > MAX_FORCINGS:
            raise ValueError(f"at most {MAX_FORCINGS} forcing terms")
        d.nforcings = len(terms)
        for i, t in enumerate(terms):
            d.forcing[i] = t.forcing_desc(ctx)
        return d


class NullPusher:
    def desc(self, ctx, restrict_mask=0):
        d = PusherDesc()
        d.kind = 0
        return d


class RestrictedPusher:
    """RestrictedPusher{T}(pusher): only species T is pushed, the others just get t += dt."""

    def __init__(self, species, pusher):
        self.species, self.pusher = species, pusher

    def desc(self, ctx, restrict_mask=0):
        return self.pusher.desc(ctx, restrict_mask=1 << self.species)
