"""Collision-process types of the reference, host side.

Each class keeps the reference's name and constructor arguments and provides
  * `totalcs(E)`  — the total cross-section used to BUILD the rate tables (host init, numpy,
    vectorised; the reference evaluates the same closed forms in Julia at table-build time), and
  * `desc(lib)`   — the flat `ptl_process_desc` handed to the device library, which owns the
    per-event `collide` samplers (csrc/physics.cuh).

Energies are in joule, cross-sections in m^2, exactly as in the reference."""
import math
import os
import re
import numpy as np

from . import constants as co

# process kind ids — must match include/particulator_b200.h
PROC_NULL, PROC_COULOMB, PROC_RBEB, PROC_MOLLER, PROC_BHABA, PROC_SELTZER, PROC_COMPTON, \
    PROC_PHOTOELECTRIC, PROC_BETHE_HEITLER, PROC_ANIHILATION, PROC_LX_EXCITATION, \
    PROC_LX_IONIZATION, PROC_LX_ATTACHMENT, PROC_LX_ELASTIC = range(14)

ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON = range(4)


def speed(species, eng):
    """speed(::Type{Electron|Positron|Photon}, eng): electron.jl:47, positron.jl:34, photon.jl:46."""
    eng = np.asarray(eng, dtype=np.float64)
    if species == PHOTON:
        return np.full_like(eng, co.c)
    if species == SLOW_ELECTRON:
        return np.sqrt(2 * eng / co.electron_mass)   # lxcat.jl:33
    return co.c * np.sqrt(1 - (co.electron_mc2 / (co.electron_mc2 + eng)) ** 2)


class CollisionProcess:
    kind = PROC_NULL
    name = "NullCollision"

    def params(self):
        return []

    def aux(self):
        return -1

    def totalcs(self, eng):
        return np.zeros_like(np.asarray(eng, dtype=np.float64))

    def __repr__(self):
        return f"{self.name}({', '.join(f'{p:g}' for p in self.params())})"


class NullCollision(CollisionProcess):
    """collisions.jl:3"""


class RelativisticCoulomb(CollisionProcess):
    """relativistic_coulomb.jl:5-8; totalcs :58-75"""
    kind = PROC_COULOMB
    name = "RelativisticCoulomb"

    def __init__(self, Z):
        self.Z = Z

    def params(self):
        return [float(self.Z)]

    def totalcs(self, K):
        Z = self.Z
        K = np.asarray(K, dtype=np.float64) + 1e-4 * co.eV
        a = 1.3413 * Z ** (-1 / 3) * co.a_0
        g = 1 + K / (co.electron_mass * co.c ** 2)
        p = np.sqrt(K * (K + 2 * (co.electron_mass * co.c ** 2))) / co.c
        beta = p / (g * co.electron_mass * co.c)
        alpha = co.hbar ** 2 / (4 * p ** 2 * a ** 2)
        return (math.pi * co.r_e ** 2 * Z ** 2 / (beta ** 4 * g ** 2) *
                ((1 + alpha * beta ** 2) / (alpha * (1 + alpha)) + beta ** 2 * np.log(alpha / (1 + alpha))))


class RBEB(CollisionProcess):
    """rbeb.jl:5-14; totalcs :117-150"""
    kind = PROC_RBEB
    name = "RBEB"

    def __init__(self, B, U, N):
        self.B, self.U, self.N = B, U, N

    def params(self):
        return [self.B, self.U, float(self.N)]

    def totalcs(self, T):
        T = np.asarray(T, dtype=np.float64)
        B, U, N = self.B, self.U, self.N
        mc2 = co.electron_mc2
        with np.errstate(all="ignore"):
            t1, b1, u1 = T / mc2, B / mc2, U / mc2
            bt2 = 1 - 1 / (1 + t1) ** 2
            bb2 = 1 - 1 / (1 + b1) ** 2
            bu2 = 1 - 1 / (1 + u1) ** 2
            t = T / B
            al = co.fine_structure
            s = (2 * co.a_0 ** 2 * math.pi * al ** 4 / (b1 * (bb2 + bt2 + bu2)) *
                 (1 - 1 / t + (b1 ** 2 * (t - 1)) / (2 * (1 + t1) ** 2)
                  - ((1 + 2 * t1) * np.log(t)) / ((1 + t) * (1 + t1) ** 2)
                  - ((t ** 2 - 1) * (bt2 - np.log(bt2) + np.log(2 * b1 * (1 - bt2)))) / (2 * t ** 2)))
        return np.where(T > B, N * s, 0.0)   # `* (T > B)`: false is a strong zero in Julia


def rbeb_dsdw(W, T, B, U):
    """rbeb.jl:87-111 — differential cross-section (tests integrate it against totalcs / the sampler)."""
    mc2 = co.electron_mc2
    t1, b1, u1 = T / mc2, B / mc2, U / mc2
    bt2 = 1 - 1 / (1 + t1) ** 2
    bb2 = 1 - 1 / (1 + b1) ** 2
    bu2 = 1 - 1 / (1 + u1) ** 2
    t = T / B
    w = W / B
    al = co.fine_structure
    return ((2 * co.a_0 ** 2 * math.pi * al ** 4 / (b1 * (bb2 + bt2 + bu2))) *
            (b1 ** 2 / (1 + t1) ** 2 + 1 / (t - w) ** 2 + 1 / (1 + w) ** 2 -
             ((1 + 2 * t1) * (1 / (t - w) + 1 / (1 + w))) / ((1 + t) * (1 + t1) ** 2) +
             (1 / (t - w) ** 3 + 1 / (1 + w) ** 3) * (-bt2 - np.log(2 * b1) + np.log(bt2 / (1 - bt2)))))


# rbeb.jl:25-48 — B and U of molecular orbitals (Hwang 1996; Santos 2003 for the K shells)
N2_ORBITALS = [RBEB(0.4095e3 * co.eV, 0.6033e3 * co.eV, 4),
               RBEB(41.72 * co.eV, 71.13 * co.eV, 2),
               RBEB(21.00 * co.eV, 63.18 * co.eV, 2),
               RBEB(17.07 * co.eV, 44.30 * co.eV, 4),
               RBEB(15.58 * co.eV, 54.91 * co.eV, 2)]
O2_ORBITALS = [RBEB(0.5438e3 * co.eV, 0.7962e3 * co.eV, 4),
               RBEB(46.19 * co.eV, 79.73 * co.eV, 2),
               RBEB(29.82 * co.eV, 90.92 * co.eV, 2),
               RBEB(19.64 * co.eV, 59.89 * co.eV, 4),
               RBEB(19.79 * co.eV, 71.84 * co.eV, 2),
               RBEB(12.07 * co.eV, 84.88 * co.eV, 2)]
ORBITALS = {"N2": N2_ORBITALS, "O2": O2_ORBITALS}


class Moller(CollisionProcess):
    """moller.jl:8-11; totalcs :40-57"""
    kind = PROC_MOLLER
    name = "Moller"

    def __init__(self, Z, tcut):
        self.Z, self.tcut = Z, tcut

    def params(self):
        return [float(self.Z), self.tcut]

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        with np.errstate(all="ignore"):
            x = self.tcut / eng
            g = 1 + eng / co.electron_mc2
            b2 = (g ** 2 - 1) / g ** 2
            A = (((g - 1) ** 2 / g ** 2) * (1 / 2 - x) + 1 / x - 1 / (1 - x)
                 - ((2 * g - 1) / g ** 2) * np.log((1 - x) / x))
            v = np.maximum(0, 2 * math.pi * co.r_e ** 2 * self.Z * A / (g - 1) / b2)
        return np.where(x > 1, 0.0, v)


def bhaba_bs(y):
    """bhaba.jl:84-91"""
    return (2 - y ** 2, (1 - 2 * y) * (3 + y ** 2), (1 - 2 * y) ** 2 + (1 - 2 * y) ** 3, (1 - 2 * y) ** 3)


class Bhaba(CollisionProcess):
    """bhaba.jl:4-7; totalcs :35-49"""
    kind = PROC_BHABA
    name = "Bhaba"

    def __init__(self, Z, tcut):
        self.Z, self.tcut = Z, tcut

    def params(self):
        return [float(self.Z), self.tcut]

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        with np.errstate(all="ignore"):
            g = 1 + eng / co.electron_mc2
            beta = np.sqrt((g ** 2 - 1) / g ** 2)
            x = self.tcut / eng
            y = 1 / (g + 1)
            B1, B2, B3, B4 = bhaba_bs(y)
            A = (1 / x - 1) / beta ** 2 + B1 * np.log(x) + B2 * (1 - x) - B3 * (1 - x ** 2) / 2 + B4 * (1 - x ** 3) / 3
            v = 2 * math.pi * co.r_e ** 2 * self.Z * A / (g - 1)
            # Julia: max(0, NaN) == NaN
            return np.where(np.isnan(v), np.nan, np.maximum(0, v))


class Compton(CollisionProcess):
    """compton.jl:1-3; totalcs :59-112 (G4KleinNishinaCompton empirical fit)"""
    kind = PROC_COMPTON
    name = "Compton"

    def __init__(self, Z):
        self.Z = Z

    def params(self):
        return [float(self.Z)]

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        Z = self.Z
        barn = 1e-28
        mc2 = co.electron_mc2
        a, b, c = 20.0, 230.0, 440.0
        d1, d2, d3, d4 = 2.7965e-1 * barn, -1.8300e-1 * barn, 6.7527 * barn, -1.9798e+1 * barn
        e1, e2, e3, e4 = 1.9756e-5 * barn, -1.0205e-2 * barn, -7.3913e-2 * barn, 2.7079e-2 * barn
        f1, f2, f3, f4 = -3.9178e-7 * barn, 6.8241e-5 * barn, 6.0480e-5 * barn, 3.0274e-4 * barn
        p1Z = Z * (d1 + e1 * Z + f1 * Z ** 2)
        p2Z = Z * (d2 + e2 * Z + f2 * Z ** 2)
        p3Z = Z * (d3 + e3 * Z + f3 * Z ** 2)
        p4Z = Z * (d4 + e4 * Z + f4 * Z ** 2)
        T0 = 15.0e3 * co.eV
        if Z <= 1:
            T0 = 40.0e3 * co.eV   # compton.jl:89 uses an undefined `keV`; only reachable for hydrogen
        with np.errstate(all="ignore"):
            X = np.maximum(eng, T0) / mc2
            sigma = p1Z * np.log(1 + 2 * X) / X + (p2Z + p3Z * X + p4Z * X ** 2) / (1 + a * X + b * X ** 2 + c * X ** 3)
            dT0 = 1e3 * co.eV
            X1 = (T0 + dT0) / mc2
            sigma1 = p1Z * math.log(1.0 + 2.0 * X1) / X1 + (p2Z + p3Z * X1 + p4Z * X1 ** 2) / (1.0 + a * X1 + b * X1 ** 2 + c * X1 ** 3)
            c1 = -T0 * (sigma1 - sigma) / (sigma * dT0)
            c2 = 0.150
            if Z > 1.5:
                c2 = 0.375 - 0.0556 * math.log(Z)
            y = np.log(eng / T0)
            low = sigma * np.exp(-y * (c1 + c2 * y))
        return np.where(eng < T0, low, sigma)


class KleinNishinaCompton(Compton):
    """compton.jl:5-7; totalcs :34-47 (closed-form Klein-Nishina, L&L 4 sect. 86)"""
    name = "KleinNishinaCompton"

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        x = 2 * eng / co.electron_mc2
        sigma = (2 * math.pi * co.r_e ** 2 * (1 / x) *
                 ((1 - 4 / x - 8 / x ** 2) * np.log(1 + x) + 1 / 2 + 8 / x - 1 / (1 + x) ** 2 / 2))
        return self.Z * sigma


# --- static data for the two elements of air (values from src/static_sandia_data.jl:134-157,
#     :1366,:1409 and src/atomic_shells.jl:36-40); other Z are parsed from the reference tree when
#     it is present (see _load_static_tables) -------------------------------------------------
_SANDIA = {
    7: [(0.01, 0.1010E+05, 0.0000E+00, 0.0000E+00, 0.0000E+00),
        (0.0404, -0.3622E+03, 0.3873E+03, 0.1244E+02, -0.4452E+00),
        (0.4016, -0.2338E+04, 0.5732E+04, -0.2082E+03, 0.1482E+03),
        (0.8, -0.4940E+01, -0.8442E+02, 0.4620E+04, -0.1186E+04),
        (4.0, 0.2019E+01, -0.1249E+03, 0.4609E+04, -0.9421E+03),
        (20.0, 0.1709E-01, -0.8196E+01, 0.2345E+04, 0.1369E+05),
        (100.0, 0.1872E-02, -0.6732E+00, 0.1282E+04, 0.5700E+05),
        (500.0, 0.8122E-03, 0.8364E+00, 0.4410E+03, 0.2358E+06)],
    8: [(0.01, 9.343E+03, 9.026E+02, -2.467E+01, 1.505E-01),
        (0.02, 6.034E+03, 7.319E+02, -2.677E+01, 2.842E-01),
        (0.0483, -0.2863E+03, 0.4085E+03, 0.4436E+02, -0.1782E+01),
        (0.532, -0.7181E+02, 0.4748E+03, 0.5542E+04, -0.1363E+04),
        (4.0, 0.2745E+01, -0.1747E+03, 0.7159E+04, -0.2213E+04),
        (20.0, 0.3774E-01, -0.1559E+02, 0.4045E+04, 0.1810E+05),
        (100.0, 0.3169E-02, -0.1146E+01, 0.2194E+04, 0.9131E+05),
        (500.0, 0.1367E-02, 0.1473E+01, 0.7214E+03, 0.4048E+06)],
}
_Z_TO_A_RATIO = {7: 0.4998, 8: 0.5}
_SHELLS_EV = {7: [403.0, 37.3, 20.33, 14.53], 8: [543.1, 41.6, 28.48, 13.62]}


def _load_static_tables(Z, refdir=None):
    """Fetch Sandia rows / Z-to-A ratio / shell energies for an arbitrary Z from the reference tree
    (data only) when it is available; air (Z = 7, 8) is built in."""
    if Z in _SANDIA:
        return
    refdir = refdir or os.environ.get("PTL_REFERENCE_DIR", "/root/reference")
    f1 = os.path.join(refdir, "src", "static_sandia_data.jl")
    f2 = os.path.join(refdir, "src", "atomic_shells.jl")
    if not (os.path.exists(f1) and os.path.exists(f2)):
        raise ValueError(f"no photo-electric data for Z={Z} (only Z=7,8 are built in)")

    def strip(txt):
        return re.sub(r"#.*", "", txt)

    def array_after(txt, name):
        m = re.search(name + r"\s*=\s*\[(.*?)\]", txt, re.S)
        return m.group(1)

    t1 = strip(open(f1).read())
    rows = [tuple(float(v) for v in m.group(1).split(",") if v.strip())
            for m in re.finditer(r"\(([^()]*)\)\s*,", array_after(t1, "const SANDIA_TABLE"))]
    nint = [int(v) for v in array_after(t1, "const NUMBER_OF_INTERVALS").split(",") if v.strip()]
    z2a = [float(v) for v in array_after(t1, "const Z_TO_A_RATIO").split(",") if v.strip()]
    start = sum(nint[:Z - 1])
    _SANDIA[Z] = rows[start:start + nint[Z - 1]]
    _Z_TO_A_RATIO[Z] = z2a[Z - 1]
    t2 = strip(open(f2).read())
    nsh = [int(v) for v in array_after(t2, "const NUMBER_OF_SHELLS").split(",") if v.strip()]
    sh = [float(v) for v in array_after(t2, "const ATOMIC_SHELLS").split(",") if v.strip()]
    s0 = sum(nsh[:Z - 1])
    _SHELLS_EV[Z] = sh[s0:s0 + nsh[Z - 1]]


def binding_energies(Z):
    """atomic_shells.jl:481-486"""
    _load_static_tables(Z)
    return np.array(_SHELLS_EV[Z]) * co.eV


class PhotoElectric(CollisionProcess):
    """photo_electric.jl:7-32 (ctor: Sandia coefficients scaled to SI, per atom); totalcs :54-58"""
    kind = PROC_PHOTOELECTRIC
    name = "PhotoElectric"

    def __init__(self, Z):
        _load_static_tables(Z)
        self.Z = Z
        s = (co.kilo * co.eV) ** np.arange(1, 5)
        A = Z / _Z_TO_A_RATIO[Z]
        data = _SANDIA[Z]
        self.left_energy = np.array([d[0] for d in data]) * (co.kilo * co.eV)
        self.coeffs = np.stack([np.array(d[1:]) * s for d in data], axis=1) * (co.centi ** 2 * A / co.N_A)
        self.binding = binding_energies(Z)

    def params(self):
        b = list(self.binding[:4])
        return [float(self.Z), float(len(b))] + b + [0.0] * (4 - len(b))

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        j = np.searchsorted(self.left_energy, eng, side="right")   # searchsortedlast
        jj = np.maximum(j, 1) - 1
        with np.errstate(all="ignore"):
            tot = sum(eng ** (-float(i)) * self.coeffs[i - 1, jj] for i in range(1, 5))
        return np.where(j > 0, tot, 0.0)


class BetheHeitler(CollisionProcess):
    """bethe_heitler.jl:1-3; totalcs :28-80 (G4BetheHeitlerModel parametrisation)"""
    kind = PROC_BETHE_HEITLER
    name = "BetheHeitler"

    def __init__(self, Z):
        self.Z = Z

    def params(self):
        return [float(self.Z)]

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        Z = self.Z
        mb = 1e-34
        lim = 1.5e6 * co.eV
        a = [8.7842e+2 * mb, -1.9625e+3 * mb, 1.2949e+3 * mb, -2.0028e+2 * mb, 1.2575e+1 * mb, -2.8333e-1 * mb]
        b = [-1.0342e+1 * mb, 1.7692e+1 * mb, -8.2381 * mb, 1.3063 * mb, -9.0815e-2 * mb, 2.3586e-3 * mb]
        c = [-4.5263e+2 * mb, 1.1161e+3 * mb, -8.6749e+2 * mb, 2.1773e+2 * mb, -2.0467e+1 * mb, 6.5372e-1 * mb]
        eng1 = np.maximum(eng, lim)
        x = np.log(eng1 / co.electron_mc2)
        x2 = x * x
        x3 = x2 * x
        x4 = x3 * x
        x5 = x4 * x
        F1 = a[0] + a[1] * x + a[2] * x2 + a[3] * x3 + a[4] * x4 + a[5] * x5
        F2 = b[0] + b[1] * x + b[2] * x2 + b[3] * x3 + b[4] * x4 + b[5] * x5
        F3 = c[0] + c[1] * x + c[2] * x2 + c[3] * x3 + c[4] * x4 + c[5] * x5
        sigma = (Z + 1) * (F1 * Z + F2 * Z * Z + F3)
        low = sigma * ((eng - 2 * co.electron_mc2) / (lim - 2 * co.electron_mc2)) ** 2
        sigma = np.where(eng < lim, low, sigma)
        return np.where(eng < 2 * co.electron_mc2, 0.0, sigma)


class PositronAnihilation(CollisionProcess):
    """anihilation.jl:1-4; totalcs :25-32 (Heitler)"""
    kind = PROC_ANIHILATION
    name = "PositronAnihilation"

    def __init__(self, Z):
        self.Z = Z

    def params(self):
        return [float(self.Z)]

    def totalcs(self, eng):
        eng = np.asarray(eng, dtype=np.float64)
        with np.errstate(all="ignore"):
            g = 1 + eng / co.electron_mc2
            A = ((g ** 2 + 4 * g + 1) / (g ** 2 - 1) * np.log(g + np.sqrt(g ** 2 - 1)) - (g + 3) / np.sqrt(g ** 2 - 1)) / (g + 1)
            return self.Z * math.pi * co.r_e ** 2 * A


class SeltzerBerger(CollisionProcess):
    """seltzer.jl:9-64 — the constructor lives in seltzer.py (`SeltzerBerger.from_Z`)."""
    kind = PROC_SELTZER
    name = "SeltzerBerger"

    def __init__(self, Z, log_energy, totalcs_tab, data, synthetic=False):
        self.Z = Z
        self.log_energy = np.ascontiguousarray(log_energy, dtype=np.float64)
        self.totalcs_tab = np.ascontiguousarray(totalcs_tab, dtype=np.float64)
        self.data = np.asfortranarray(data, dtype=np.float64)     # [ncum, nE], ncum fastest
        self.synthetic = synthetic
        self._aux = -1

    def params(self):
        return [float(self.Z)]

    def aux(self):
        return self._aux

    def totalcs(self, K):
        """seltzer.jl:217-231: linear in log K between tabulated energies, 0 below the first."""
        K = np.asarray(K, dtype=np.float64)
        with np.errstate(all="ignore"):
            logK = np.log(K)
        le = self.log_energy
        i = np.searchsorted(le, logK, side="right")     # searchsortedlast (1-based)
        ii = np.clip(i, 1, len(le) - 1)
        w = (le[ii] - logK) / (le[ii] - le[ii - 1])
        v = w * self.totalcs_tab[ii - 1] + (1 - w) * self.totalcs_tab[ii]
        return np.where(i == 0, 0.0, v)


# --- LXCat process kinds (slow-electron.jl:70-84) -------------------------------------------
class Excitation(CollisionProcess):
    kind = PROC_LX_EXCITATION
    name = "Excitation"

    def __init__(self, threshold):
        self.threshold = threshold

    def params(self):
        return [self.threshold]


class Ionization(Excitation):
    kind = PROC_LX_IONIZATION
    name = "Ionization"


class Attachment(Excitation):
    kind = PROC_LX_ATTACHMENT
    name = "Attachment"


class Elastic(CollisionProcess):
    kind = PROC_LX_ELASTIC
    name = "Elastic"

    def __init__(self, mass_ratio):
        self.mass_ratio = mass_ratio

    def params(self):
        return [self.mass_ratio]
