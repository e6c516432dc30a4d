"""BASELINE config 4: mixed e-/gamma/e+ population (feedback regime) — Compton, photo-electric, pair production,
annihilation, Bhabha, bremsstrahlung all active.  Device-timed advance! + droplow! per step; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import particulator_b200 as P
co = P.co


def directions(rng, n, cmin=-1.0):
    cost = rng.uniform(cmin, 1, n); phi = rng.uniform(0, 2 * np.pi, n); sint = np.sqrt(1 - cost ** 2)
    return np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)


def main(ne=10_000_000, ng=10_000_000, npos=200_000, steps=4, warmup=2, dt=2.5e-11):
    comp = P.air_composition()
    Fdt = co.elementary_charge * 5e5 * dt
    tabs = {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}
    ctx = P.Context(device=0, stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(3)
    Ke = np.clip(rng.exponential(7.3e6, ne), 1e3, 1e8) * co.eV
    Kg = np.exp(rng.uniform(np.log(1e4), np.log(3e7), ng)) * co.eV                 # dN/dE ~ 1/E on [10 keV, 30 MeV]
    Kp = np.exp(rng.uniform(np.log(1e5), np.log(2e7), npos)) * co.eV
    def st(sp, K, d):
        return dict(x=np.zeros((len(K), 3)), p=d * P.momentum_norm_from_kin(sp, K)[:, None], s=-np.log(1 - rng.random(len(K))))
    el = P.Population(ctx, P.ELECTRON, int(1.8 * ne), st(P.ELECTRON, Ke, directions(rng, ne, 0.8)), tabs["electron"], 1e3 * co.eV)
    ph = P.Population(ctx, P.PHOTON, int(1.5 * ng), st(P.PHOTON, Kg, directions(rng, ng)), tabs["photon"], 1e3 * co.eV)
    po = P.Population(ctx, P.POSITRON, int(4 * npos) + (1 << 18), st(P.POSITRON, Kp, directions(rng, npos)), tabs["positron"], 1e2 * co.eV)
    mp = P.MultiPopulation(("electron", el), ("photon", ph), ("positron", po))
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0, 0, -5e5]), P.HomogeneousField([0, 0, 0])))
    t = 0.0; ms_tot = 0.0; rows = sub = 0; per = []
    for it in range(warmup + steps):
        n0 = [len(q) for q in (el, ph, po)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t += dt
        P.advance(mp, psh, t)
        stt = P.last_advance_stats(mp)
        for q in mp: P.droplow(q)
        e1.record(); torch.cuda.synchronize()
        if it >= warmup:
            ms = e0.elapsed_time(e1); ms_tot += ms; rows += sum(n0); sub += stt["substeps"]; per.append(round(ms, 2))
    print(json.dumps({"workload": "mixed e-/gamma/e+ (BASELINE configs[3])", "n_start": [ne, ng, npos], "n_end": [len(q) for q in (el, ph, po)],
                      "steps": steps, "ms_per_step": ms_tot / steps, "step_ms": per, "particle_steps_per_s": rows / (ms_tot * 1e-3),
                      "substeps_per_s": sub / (ms_tot * 1e-3), "kappa_all_species": sub / rows, "passes_last": stt["passes"], "flags": ctx.error_flags()}))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--ne", type=int, default=10_000_000); ap.add_argument("--ng", type=int, default=10_000_000)
    ap.add_argument("--npos", type=int, default=200_000); ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    main(a.ne, a.ng, a.npos, a.steps)
