"""Seltzer-Berger bremsstrahlung tables — host-side construction.

Mirrors the reference constructor `SeltzerBerger(d::RawG4Physics2DVector, Z; ...)`
(src/seltzer.jl:23-49), `findcumvalues!` (:163-192), `scaledcs` (:201-211) and the Geant4 2-D
vector reader (src/util.jl:165-190).  BSplineKit's order-2 interpolant and its integral
(seltzer.jl:164-174) are restated in closed form: a piecewise-linear interpolant of the scaled
DCS against ln(k/T) and its exact piecewise-quadratic antiderivative.

The raw Geant4 `br{Z}` files carry a "use within Geant4 / non-commercial" notice
(data/brem_SB/README) and are therefore NOT vendored: they are read from the reference tree (or
`$PTL_DATA_DIR`) when present and the derived tables are cached under build/ (git-ignored, but
shipped to the GPU box).  When neither exists a synthetic same-shape table is generated so that
benchmarks never depend on the data files; `SeltzerBerger.synthetic` records which one was used."""
import os
import math
import numpy as np

from . import constants as co
from .processes import SeltzerBerger

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = os.path.join(_ROOT, "build", "tables")


class RawG4Physics2DVector:
    """util.jl:165-190: line 1 `k nx ny`, line 2 x[nx], line 3 y[ny], then value[ny, nx]."""

    def __init__(self, fname):
        with open(fname) as f:
            self.k, self.nx, self.ny = (int(v) for v in f.readline().split())
            self.x = np.array([float(v) for v in f.readline().split()])
            self.y = np.array([float(v) for v in f.readline().split()])
            self.value = np.loadtxt(f)
        assert len(self.x) == self.nx and len(self.y) == self.ny
        assert self.value.shape == (self.ny, self.nx)


class _LinearCumInt:
    """Piecewise-linear interpolant p(x) through (x_i, p_i) and its exact integral from x_0."""

    def __init__(self, x, p):
        self.x = np.asarray(x, dtype=np.float64)
        self.p = np.asarray(p, dtype=np.float64)
        dx = np.diff(self.x)
        self.cum = np.concatenate([[0.0], np.cumsum(0.5 * (self.p[1:] + self.p[:-1]) * dx)])

    def _seg(self, xq):
        return np.clip(np.searchsorted(self.x, xq, side="right") - 1, 0, len(self.x) - 2)

    def density(self, xq):
        i = self._seg(xq)
        t = (xq - self.x[i]) / (self.x[i + 1] - self.x[i])
        return self.p[i] + t * (self.p[i + 1] - self.p[i])

    def __call__(self, xq):
        i = self._seg(xq)
        h = xq - self.x[i]
        slope = (self.p[i + 1] - self.p[i]) / (self.x[i + 1] - self.x[i])
        return self.cum[i] + self.p[i] * h + 0.5 * slope * h * h


def findcumvalues(x, p, pcum, xmin, xmax, rtol=1e-6):
    """seltzer.jl:163-195: x-values at which the normalised cumulative of p reaches pcum."""
    ci = _LinearCumInt(x, p)
    assert np.all(ci.p > 0)
    cum0 = float(ci(max(xmin, x.min())))
    cum1 = float(ci(min(xmax, x.max())))
    # knots of the integral spline: data sites with the end points repeated (order 3)
    knt = np.concatenate([[x[0], x[0]], x, [x[-1], x[-1]]])
    fknt = (ci(knt) - cum0) / (cum1 - cum0)
    out = np.empty(len(pcum))
    for i, pc in enumerate(pcum):
        j = int(np.searchsorted(fknt, pc, side="right"))      # searchsortedlast, 1-based
        lo = knt[max(1, j) - 1]
        hi = knt[min(len(knt), j + 1) - 1]
        xsol = 0.5 * (lo + hi)
        dx = math.inf
        it = 0
        while xsol != 0.0 and abs(dx / xsol) > rtol:     # Julia: abs(Inf/0)=Inf loops, abs(x/0)=NaN/Inf at xsol == 0 stops only via NaN; the k/T = 1 end point is exact
            f = (float(ci(xsol)) - cum0) / (cum1 - cum0)
            df = float(ci.density(xsol)) / (cum1 - cum0)
            dx = (f - pc) / df
            xsol = xsol - dx
            it += 1
            if it > 200:
                raise RuntimeError("findcumvalues: Newton iteration did not converge")
        out[i] = xsol
    return out


def scaledcs(logk, s, logkmin, logkmax):
    """seltzer.jl:201-211"""
    ci = _LinearCumInt(logk, s)
    return float(ci(min(logkmax, logk.max()))) - float(ci(max(logkmin, logk.min())))


def build_from_raw(d, Z, ncum=1000, gamma_min=1e2 * co.eV, gamma_max=5e7 * co.eV, energy_scale=1e6 * co.eV,
                   synthetic=False):
    """seltzer.jl:23-49"""
    pcum = np.linspace(0.0, 1.0, ncum)
    log_energy = np.log(np.exp(d.y) * energy_scale)
    data = np.zeros((ncum, len(log_energy)))
    totalcs = np.zeros(len(log_energy))
    mc2 = co.electron_mc2
    logk = np.log(d.x)
    for i in range(len(log_energy)):
        T1 = math.exp(log_energy[i])
        beta = math.sqrt(1 - 1 / (1 + T1 / mc2) ** 2)
        data[:, i] = findcumvalues(logk, d.value[i, :], pcum, math.log(gamma_min / T1), math.log(gamma_max / T1))
        totalcs[i] = ((Z ** 2 / beta ** 2) * 1e-31 *
                      scaledcs(logk, d.value[i, :], math.log(gamma_min / T1), math.log(gamma_max / T1)))
    return SeltzerBerger(Z, log_energy, totalcs, data, synthetic=synthetic)


class _SyntheticRaw:
    """Same-shape stand-in for a Geant4 `br{Z}` vector: 32 k/T fractions x 57 energies with a smooth,
    positive scaled DCS  chi(k/T, T) ~ (a + b(1 - k/T)^2) * (1 + c ln(1 + T/mc2)) [mb], which has the
    right order of magnitude for light elements.  Used only when the data files are unavailable."""

    def __init__(self, Z):
        self.k, self.nx, self.ny = 4, 32, 57
        self.x = np.array([1e-12, 0.025, 0.05, 0.075, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.55,
                           0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.925, 0.95, 0.97, 0.99, 0.995, 0.999,
                           0.9995, 0.9999, 0.99995, 0.99999, 1.0])
        e_kev = []
        for dec in range(0, 7):
            for m in (1.0, 1.5, 2.0, 3.0, 4.0, 5.0, 6.0, 8.0):
                e_kev.append(m * 10.0 ** dec)
        e_kev = np.array(e_kev[:56] + [1e7])
        e_kev[0] = 0.99995                     # like the Geant4 grid, start just below 1 keV
        self.y = np.log(e_kev * 1e-3)          # ln(T / MeV)
        T = e_kev * 1e3 * co.eV
        tau = T / co.electron_mc2
        f = self.x[None, :]
        self.value = (4.0 + 6.0 * (1 - f) ** 2) * (1 + 0.35 * np.log1p(tau))[:, None] * (1 + 0.02 * (Z - 7))


def _raw_path(Z):
    for base in (os.environ.get("PTL_DATA_DIR"), os.path.join(os.environ.get("PTL_REFERENCE_DIR", "/root/reference"), "data")):
        if base:
            p = os.path.join(base, "brem_SB", f"br{Z}")
            if os.path.exists(p):
                return p
    return None


def from_Z(Z, allow_synthetic=True, use_cache=True, **kw):
    """SeltzerBerger(Z) (seltzer.jl:56-63).  Order of preference: cached derived table under
    build/tables, the raw Geant4 file from the reference tree, a synthetic same-shape table."""
    cache = os.path.join(_CACHE, f"sb_Z{Z}.npz")
    if use_cache and not kw and os.path.exists(cache):
        z = np.load(cache)
        return SeltzerBerger(Z, z["log_energy"], z["totalcs"], z["data"], synthetic=bool(z["synthetic"]))
    raw = _raw_path(Z)
    if raw is not None:
        sb = build_from_raw(RawG4Physics2DVector(raw), Z, **kw)
    elif allow_synthetic:
        sb = build_from_raw(_SyntheticRaw(Z), Z, synthetic=True, **kw)
    else:
        raise FileNotFoundError(f"brem_SB/br{Z} not found (set PTL_DATA_DIR)")
    if use_cache and not kw:
        try:
            os.makedirs(_CACHE, exist_ok=True)
            np.savez(cache, log_energy=sb.log_energy, totalcs=sb.totalcs_tab, data=sb.data, synthetic=sb.synthetic)
        except OSError:
            pass
    return sb
