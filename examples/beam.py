"""The monoenergetic-beam configuration of the reference's scripts/beam.jl (BASELINE configs[1]) written against this
package's host API: same parameters (7 MeV electrons along +z with the 1e-6 p_y offset that keeps `turn` off its pole, air at
STP, E = 5e5 V/m between z1 and z2, dt = 2.5e-11 s, cuts 1 keV / 1 keV / 100 eV, safety 1.15), same call sequence
(tables -> populations -> MultiPopulation -> init! -> run!), optional walls on z.  Prints the populations and the energy
spectrum of the electrons at the end.

    python examples/beam.py --n 1000 --tfinal 1e-8 [--zwall 1.0]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import particulator_b200 as P

co = P.co


def main(n_init_particles=1, maxp=None, init_energy=7e6 * co.eV, dt=2.5e-11, efield=5e5, safety=1.15, tfinal=1e-8, z1=0.0, z2=200.0,
         seed=0, Kthresh=1e3 * co.eV, output_dt=None, zwall=None, device=0, verbosity=0, ctx=None):
    maxp = maxp or max(1_000_000, 10 * n_init_particles)
    composition = P.air_composition()                       # N2 0.79 / O2 0.21 at co.nair
    Fdt = co.elementary_charge * efield * dt
    ecolls = P.build_electron_collision_table(composition, Fdt, safety=safety)
    pcolls = P.build_positron_collision_table(composition, 1e2 * co.eV, Fdt, safety=safety)
    gcolls = P.build_photon_collision_table(composition)

    if ctx is None:
        ctx = P.Context(device=device)                      # no B200, no run: there is no CPU fallback (tests pass the oracle's context)
    ctx.set_rng(seed, 0)
    pnorm = P.momentum_norm_from_kin(P.ELECTRON, np.array([init_energy]))[0]
    p0 = np.tile([0.0, 1e-6 * pnorm, pnorm], (n_init_particles, 1))
    init = dict(x=np.zeros((n_init_particles, 3)), p=p0)
    electrons = P.Population(ctx, P.ELECTRON, maxp, init, ecolls, Kthresh, rng=np.random.default_rng(seed))
    photons = P.Population(ctx, P.PHOTON, maxp, None, gcolls, 1e3 * co.eV)
    positrons = P.Population(ctx, P.POSITRON, maxp, None, pcolls, 1e2 * co.eV)
    mpopl = P.MultiPopulation(("electron", electrons), ("photon", photons), ("positron", positrons))
    P.init(mpopl)

    pusher = P.RK2Pusher(P.ElectromagneticField(P.DoubleLayerField(z1, z2, [0.0, 0.0, -efield]), P.HomogeneousField([0.0, 0.0, 0.0])))
    if zwall is not None:
        callback = P.CombinedCallback([P.WallCallback(P.ELECTRON, 3, zwall), P.WallCallback(P.PHOTON, 3, zwall),
                                       P.WallCallback(P.POSITRON, 3, zwall)])
    else:
        callback = P.VoidCallback()
    t = P.run(mpopl, pusher, tfinal, dt, callback, output_dt=output_dt, verbosity=verbosity)
    return dict(ctx=ctx, mpopl=mpopl, electrons=electrons, photons=photons, positrons=positrons, callback=callback, t=t)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--tfinal", type=float, default=1e-8)
    ap.add_argument("--zwall", type=float, default=None)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import time
    t0 = time.time()
    r = main(n_init_particles=a.n, tfinal=a.tfinal, zwall=a.zwall, seed=a.seed)
    el, ph, po = r["electrons"], r["photons"], r["positrons"]
    print(f"t = {r['t'] / 1e-9:.2f} ns after {time.time() - t0:.1f} s wall: electrons {len(el)}, photons {len(ph)}, positrons {len(po)}")
    print(f"electrons: mean energy {P.meanenergy(el) / (1e6 * co.eV):.3f} MeV, max {P.maxenergy(el) / (1e6 * co.eV):.3f} MeV, "
          f"centroid z = {P.spread(el)[0][2]:.3f} m")
    h = el.histogram("energy", 1e3 * co.eV, 1e8 * co.eV, 10, logscale=True)
    print("electron spectrum, counts per half-decade bin from 1 keV to 100 MeV:", [int(v) for v in h])
