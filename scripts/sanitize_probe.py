"""Workload for compute-sanitizer (racecheck / memcheck / synccheck): one advance! + droplow! of a three-species population
through a chosen lepton kernel variant.  No torch: the sanitizer only sees the library's own kernels.

  compute-sanitizer --tool racecheck python scripts/sanitize_probe.py --kernel 5 --n 50000
kernel: 3 = bq (list-scheduled), 4 = wf (re-sorting), 5 = wq (warp-private pools); --small-pass 1000000 sends every lepton
pass to the one-particle-per-lane kernel; photons always exercise the streaming kernel + its deferred-row pass."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import particulator_b200 as P
from conftest import make_world, default_pusher

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", type=int, default=5)
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--small-pass", type=int, default=0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--dt-scale", type=float, default=1.0, help="scale of dt = 2.5e-11 s; 1/2048 gives kappa ~ 1, so that from the second "
                "advance! on the leptons take the streaming kernels")
ap.add_argument("--stream-tma", type=int, default=0, help="1: the TMA-staged lepton streaming kernel")
a = ap.parse_args()
co = P.co
comp = P.air_composition()
Fdt = co.elementary_charge * 5e5 * 2.5e-11
tables = {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
          "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
          "photon": P.build_photon_collision_table(comp)}
ctx = P.Context(device=0)
ctx.set_option("kernel", a.kernel)
ctx.set_option("small_pass_rows", a.small_pass)
ctx.set_option("stream_tma", a.stream_tma)
ctx.set_rng(1, 0)
mp, el, ph, po = make_world(ctx, tables, a.n, a.n // 2, a.n // 10, cap=4 * a.n, seed=3)
t = 0.0
for _ in range(a.steps):
    t += 2.5e-11 * a.dt_scale
    P.advance(mp, default_pusher(), t)
    for q in (el, ph, po):
        P.droplow(q)
st = P.last_advance_stats(mp)
print(f"kernel {a.kernel} small_pass {a.small_pass} dt_scale {a.dt_scale:g} stream_tma {a.stream_tma}: n = {[len(q) for q in (el, ph, po)]}, substeps {st['substeps']}, passes {st['passes']}, flags {ctx.error_flags()}")
ctx.close()
