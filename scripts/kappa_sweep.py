"""kappa-sweep: the same RREA electron population advanced with dt scaled so that the mean
number of collision sub-steps per particle-step kappa spans ~0.1 ... 150; reports particle-steps/s, sub-steps/s and
the algorithmic HBM fraction (162 B per particle-step) for each dt.  One JSON line per dt."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import particulator_b200 as P
import bench

co = P.co


def run(dt, n, steps=4, warmup=3, spectrum_emin=1e3):
    comp = P.air_composition()
    Fdt = co.elementary_charge * bench.EFIELD * dt
    tabs = {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}
    ctx = P.Context(device=0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_profiling(True)
    mp, el, ph, po = bench.make_world(P, ctx, tabs, int(3.0 * n) + 4096, max(n, 1 << 20), 1 << 18)
    bench.synth_electrons_device(torch, P, el, n, seed=7, uid0=1)
    psh = bench.pusher(P)
    t = 0.0
    tot_ps = tot_sub = 0
    ms = 0.0
    kern_ms = kern_rows = 0.0
    for it in range(warmup + steps):
        n0 = len(el)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t += dt
        P.advance(mp, psh, t)
        st = P.last_advance_stats(mp)
        for q in mp:
            P.droplow(q)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            tot_ps += n0; tot_sub += st["substeps"]; ms += e0.elapsed_time(e1)
            kern_ms += st["main_ms"]; kern_rows += st["main_rows"]
    peak, _ = bench.measured_hbm_peak()
    out = {"dt_s": dt, "n": n, "kappa": tot_sub / tot_ps, "particle_steps_per_s": tot_ps / (ms * 1e-3), "substeps_per_s": tot_sub / (ms * 1e-3),
           "ms_per_step": ms / steps, "algorithmic_GBps_step": 162 * tot_ps / (ms * 1e-3) / 1e9,
           "hbm_frac_step": 162 * tot_ps / (ms * 1e-3) / 1e9 / peak,
           "main_kernel_ms": kern_ms / steps, "hbm_frac_main_kernel": (162 * kern_rows / (kern_ms * 1e-3) / 1e9 / peak) if kern_ms else None}
    ctx.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    scales = (1 / 2048,) if len(sys.argv) > 2 and sys.argv[2] == "one" else (1 / 2048, 1 / 256, 1 / 32, 1 / 4, 1.0, 4.0)
    for scale in scales:
        print(json.dumps(run(2.5e-11 * scale, n)), flush=True)
