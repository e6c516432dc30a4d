"""The avalanche-with-population-control configuration of the reference's scripts/swarm.jl (BASELINE configs[0]) on this
package's host API: one 7 MeV seed electron in air under E = 5e5 V/m, advanced in outer iterations of 1 ns (40 steps of
dt = 2.5e-11 s); after each iteration the electron population is rouletted back to `ntarget` (weights grow by 1/p, so the
weighted count keeps following the avalanche).  Everything per particle — advance!, droplow!, nactives, roulette! — runs on
the device.  The reference calls `run!(..., t + tstep, ...)`, which restarts its clock at 0 every time; the inner loop here is
the body of run! (advance! to t + dt, droplow! every population) continued from the current time.

    python examples/swarm.py --iterations 30 --ntarget 10000
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import particulator_b200 as P

co = P.co


def main(n_init_particles=1, maxp=1_000_000, ntarget=10000, init_energy=7e6 * co.eV, dt=2.5e-11, efield=5e5, safety=1.15, z1=0.0,
         z2=200.0, seed=0, Kthresh=1e3 * co.eV, iterations=300, tstep=1e-9, device=0, ctx=None, log=None):
    composition = P.air_composition()
    Fdt = co.elementary_charge * efield * dt
    ecolls = P.build_electron_collision_table(composition, Fdt, safety=safety)
    pcolls = P.build_positron_collision_table(composition, Kthresh, Fdt, safety=safety)
    gcolls = P.build_photon_collision_table(composition)
    if ctx is None:
        ctx = P.Context(device=device)
    ctx.set_rng(seed, 0)
    pnorm = P.momentum_norm_from_kin(P.ELECTRON, np.array([init_energy]))[0]
    init = dict(x=np.zeros((n_init_particles, 3)), p=np.tile([0.0, 1e-6 * pnorm, pnorm], (n_init_particles, 1)))
    electrons = P.Population(ctx, P.ELECTRON, maxp, init, ecolls, Kthresh, rng=np.random.default_rng(seed))
    photons = P.Population(ctx, P.PHOTON, maxp, None, gcolls, Kthresh)
    positrons = P.Population(ctx, P.POSITRON, maxp, None, pcolls, Kthresh)
    mpopl = P.MultiPopulation(("electron", electrons), ("photon", photons), ("positron", positrons))
    P.init(mpopl)
    pusher = P.RK2Pusher(P.ElectromagneticField(P.DoubleLayerField(z1, z2, [0.0, 0.0, -efield]), P.HomogeneousField([0.0, 0.0, 0.0])))
    t = 0.0
    history = []
    nsteps = int(round(tstep / dt))
    for i in range(iterations):
        for _ in range(nsteps):                      # body of run! (run.jl:6-9)
            t += dt
            P.advance(mpopl, pusher, t)
            for popl in mpopl:
                P.droplow(popl)
        n = P.nactives(electrons)
        if n > ntarget:
            P.roulette(ntarget / n, electrons)
            P.repack(electrons)
        history.append((t, n, P.weight(electrons)))
        if log:
            log(f"t = {t / 1e-9:6.1f} ns  electrons {n:8d}  weighted {P.weight(electrons):12.4g}  photons {len(photons)}")
    return dict(ctx=ctx, mpopl=mpopl, electrons=electrons, photons=photons, positrons=positrons, t=t, history=history)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--ntarget", type=int, default=10000)
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    r = main(n_init_particles=a.n, ntarget=a.ntarget, iterations=a.iterations, seed=a.seed, log=print)
    h = r["history"]
    if len(h) > 10 and h[-1][2] > 0 and h[len(h) // 2][2] > 0:
        rate = np.log(h[-1][2] / h[len(h) // 2][2]) / (h[-1][0] - h[len(h) // 2][0])
        print(f"avalanche growth rate over the second half: {rate:.3e} 1/s  (e-folding time {1e9 / rate:.1f} ns)")
