"""Quick device-timed probe of the advance kernel (development aid; bench.py is the contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import particulator_b200 as P
co = P.co

def main_slow(n, steps, nprocs, dt=1e-12, grid_kind=0):
    """BASELINE config 5: LXCat-style slow-electron swarm (linear table, explicit null row, null-collision dominated)."""
    tab = P.synthetic_lxcat_table(grid_kind=grid_kind, extra_levels=max(0, nprocs - 7))   # a real N2/O2 set has 50-80 channels
    stream = torch.cuda.current_stream().cuda_stream
    ctx = P.Context(device=0, stream=stream)
    rng = np.random.default_rng(0)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * co.eV / co.electron_mass) * 1.2, s=-np.log(1 - rng.random(n)))
    pop = P.Population(ctx, P.SLOW_ELECTRON, int(1.5 * n), st, tab, 0.0)
    mp = P.MultiPopulation(("slow", pop))
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -100 * co.Td * co.nair]), None))
    t = 0.0
    for it in range(steps):
        n0 = len(pop)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t += dt
        P.advance(mp, psh, t)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); stt = P.last_advance_stats(mp)
        P.droplow(pop)
        print(f"slow n={n0} procs={len(tab.proc)} step {it}: advance {ms:.2f} ms substeps={stt['substeps']} kappa={stt['substeps']/max(stt['rows'],1):.1f} births={stt['births']} "
              f"-> {n0/ms*1e3:.3e} particle-steps/s, {stt['substeps']/ms*1e3:.3e} substeps/s", flush=True)


def main(n=2_000_000, species="electron", steps=3, emin=1e3, emax=1e8, spectrum="exp"):
    comp = P.air_composition(); dt = 2.5e-11
    Fdt = co.elementary_charge * 5e5 * dt
    tabs = {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}
    stream = torch.cuda.current_stream().cuda_stream
    ctx = P.Context(device=0, stream=stream)
    rng = np.random.default_rng(0)
    sp = {"electron": P.ELECTRON, "photon": P.PHOTON, "positron": P.POSITRON}[species]
    if spectrum == "exp":
        K = np.clip(rng.exponential(7.3e6, n), emin, emax) * co.eV
    else:
        K = np.exp(rng.uniform(np.log(emin), np.log(emax), n)) * co.eV
    pn = P.momentum_norm_from_kin(sp, K)
    cost = rng.uniform(0.8, 1, n); phi = rng.uniform(0, 2 * np.pi, n); sint = np.sqrt(1 - cost ** 2)
    d = np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)
    st = dict(x=np.zeros((n, 3)), p=d * pn[:, None], s=-np.log(1 - rng.random(n)))
    cap = int(1.5 * n)
    pops = {"electron": P.Population(ctx, P.ELECTRON, cap, st if species == "electron" else None, tabs["electron"], 1e3 * co.eV),
            "photon": P.Population(ctx, P.PHOTON, cap, st if species == "photon" else None, tabs["photon"], 1e3 * co.eV),
            "positron": P.Population(ctx, P.POSITRON, cap, st if species == "positron" else None, tabs["positron"], 1e2 * co.eV)}
    mp = P.MultiPopulation(*pops.items())
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0, 0, -5e5]), P.HomogeneousField([0, 0, 0])))
    t = 0.0
    ctx.set_profiling(True)
    for it in range(steps):
        n0 = len(pops[species])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t += dt
        P.advance(mp, psh, t)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        stt = P.last_advance_stats(mp)
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for q in mp: P.droplow(q)
        e3.record(); torch.cuda.synchronize()
        print(f"{species} n={n0} step {it}: advance {ms:.2f} ms, droplow {e2.elapsed_time(e3):.2f} ms, substeps={stt['substeps']} kappa={stt['substeps']/max(stt['rows'],1):.1f} "
              f"births={stt['births']} passes={stt['passes']} main_ms={stt['main_ms']:.2f} -> {n0/ms*1e3:.3e} particle-steps/s, {stt['substeps']/ms*1e3:.3e} substeps/s, "
              f"{n0*162/ms*1e-6:.1f} GB/s algorithmic", flush=True)

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000); ap.add_argument("--species", default="electron")
    ap.add_argument("--steps", type=int, default=3); ap.add_argument("--spectrum", default="exp")
    ap.add_argument("--emin", type=float, default=1e3); ap.add_argument("--emax", type=float, default=1e8)
    ap.add_argument("--lx-procs", type=int, default=7); ap.add_argument("--dt", type=float, default=1e-12)
    a = ap.parse_args()
    if a.species == "slow":
        main_slow(a.n, a.steps, a.lx_procs, a.dt)
    else:
        main(a.n, a.species, a.steps, a.emin, a.emax, a.spectrum)
