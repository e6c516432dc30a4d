#!/usr/bin/env python
"""bench.py — particle-steps/s of the advance hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[2]): relativistic runaway electron avalanche,
electrons with E ~ exp(-E/7.3 MeV) on [1 keV, 100 MeV], cos(theta_z) ~ U[0.8, 1], uniform E_z = -5e5 V/m, STP air
(Coulomb + Seltzer-Berger + RBEB tables of scripts/beam.jl:94-105, safety 1.15), dt = 2.5e-11 s, photon and
positron populations start empty.  One "step" = what run! does per dt (src/run.jl:6-9): advance!(mpopl, pusher, t+dt)
followed by droplow! on every population.  A particle-step = one particle alive at the start of a step carried
through it.  Per-GPU work is fixed (weak scaling): N GPUs hold N x n_per_gpu electrons, sharded with no
data-path collective; every rank uses the SAME seed (streams are keyed by particle uid, so results do not depend on the
number of GPUs).  At N > 1 a second, strong-scaling measurement runs the named 1e8-electron avalanche split over the N
GPUs with the library's NCCL collectives inside the timed region (`strong`): ptl_diag_allreduce every step (what run!
prints, src/run.jl:31-40) and ptl_rebalance every second step.

  value     whole-job particle-steps/s, state resident in HBM (timed with CUDA events, max over ranks)
  e2e       the same through the C-ABI with HOST buffers: pinned host arrays -> ptl_population_upload ->
            ptl_advance -> ptl_droplow -> ptl_population_download, copies inside the timed region
  roofline  dominant kernel (electron advance, first pass): algorithmic bytes = 162 B per particle-step
             / its cudaEvent duration, against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, OpenMP over all host cores) on a bounded
            sample of the same workload — a reported baseline, not the target
  secondary the other BASELINE.json configurations and the latency regime, a few steps each (N = 1 only): photon
            streaming, kappa ~ 1 electrons, the mixed e-/gamma/e+ population (configs[3]), the LXCat slow-electron swarm
            with 64 channels (configs[4]), and the step latency of a 1e4-electron swarm (configs[0] size)
`--impl reference` times that CPU restatement alone (Julia is not installable in this image)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 2.5e-11
EFIELD = 5e5
ALGO_BYTES_PER_PARTICLE_STEP = 162.0     # 81 B read + 81 B write


def build_tables(P):
    co = P.co
    comp = P.air_composition()
    Fdt = co.elementary_charge * EFIELD * DT
    return {"electron": P.build_electron_collision_table(comp, Fdt, safety=1.15),
            "positron": P.build_positron_collision_table(comp, 1e2 * co.eV, Fdt, safety=1.15),
            "photon": P.build_photon_collision_table(comp)}


def synth_electrons_numpy(P, n, seed, uid0):
    """RREA spectrum, host arrays (used by the CPU legs and as the pinned e2e source)."""
    co = P.co
    rng = np.random.default_rng(seed)
    K = np.clip(rng.exponential(7.3e6, n), 1e3 * 1.0001, 1e8) * co.eV
    pn = P.momentum_norm_from_kin(P.ELECTRON, K)
    cost = rng.uniform(0.8, 1.0, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    sint = np.sqrt(1 - cost ** 2)
    p = np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1) * pn[:, None]
    return dict(x=np.zeros((n, 3)), p=p, w=np.ones(n), t=np.zeros(n), s=-np.log(1 - rng.random(n)), r=np.zeros(n),
                active=np.ones(n, dtype=np.uint8), uid=np.arange(uid0, uid0 + n, dtype=np.uint64))


class _DevArray:
    """Zero-copy view of a library-owned device column for torch (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def column_tensor(torch, pop, col, n):
    typestr = "<f8" if col < 10 else ("|u1" if col == 10 else "<i8")      # uid viewed as int64 (same bits)
    return torch.as_tensor(_DevArray(pop.column_ptr(col), n, typestr), device="cuda")


def synth_electrons_device(torch, P, pop, n, seed, uid0):
    """Same distribution generated on the device straight into the library's SoA columns."""
    co = P.co
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    u = lambda: torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
    K = torch.clamp(-7.3e6 * torch.log1p(-u()), 1e3 * 1.0001, 1e8) * co.eV
    pn = torch.sqrt((K + co.electron_mc2) ** 2 - co.electron_mc2 ** 2) / co.c
    cost = 0.8 + 0.2 * u()
    phi = 2 * np.pi * u()
    sint = torch.sqrt(1 - cost ** 2)
    cols = [torch.zeros(n, device="cuda", dtype=torch.float64)] * 3 + [pn * sint * torch.cos(phi), pn * sint * torch.sin(phi), pn * cost]
    cols += [torch.ones(n, device="cuda", dtype=torch.float64), torch.zeros(n, device="cuda", dtype=torch.float64),
             -torch.log1p(-u()), torch.zeros(n, device="cuda", dtype=torch.float64)]
    for c, v in enumerate(cols):
        column_tensor(torch, pop, c, n).copy_(v)
    column_tensor(torch, pop, 10, n).fill_(1)
    column_tensor(torch, pop, 11, n).copy_(torch.arange(n, device="cuda", dtype=torch.int64) + uid0)
    torch.cuda.synchronize()
    pop.set_n(n)


def make_world(P, ctx, tables, cap_e, cap_g, cap_p):
    co = P.co
    el = P.Population(ctx, P.ELECTRON, cap_e, None, tables["electron"], 1e3 * co.eV)
    ph = P.Population(ctx, P.PHOTON, cap_g, None, tables["photon"], 1e3 * co.eV)
    po = P.Population(ctx, P.POSITRON, cap_p, None, tables["positron"], 1e2 * co.eV)
    mp = P.MultiPopulation(("electron", el), ("photon", ph), ("positron", po))
    return mp, el, ph, po


def pusher(P):
    return P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -EFIELD]), P.HomogeneousField([0.0, 0.0, 0.0])))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): ONE
    `nvidia-smi -lms 200` process started before and killed after, as the recipe says.  (Spawning a fresh nvidia-smi every
    200 ms re-initialises NVML each time and stalled the CUDA calls of the pass loop: the non-kernel part of a step varied
    between 100 and 290 ms from box to box.)"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def _collect(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            try:
                self.proc.kill()
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
        import datetime
        for line in (out or "").splitlines():
            if not line.strip():
                continue
            r = [v.strip() for v in line.split(",")]
            try:                   # timestamp: YYYY/MM/DD HH:MM:SS.mmm (local time)
                ts = datetime.datetime.strptime(r[-1], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = None
            if ts is None or ts >= self.t_begin - 0.05:
                self.rows.append(r)

    def stop(self):
        self._collect()
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def shard_bounds(n, nshards, nworkers, ramp):
    """Row boundaries of the e2e shards.  Shard sizes ramp up over the first `nworkers` shards and down over the last ones
    (pipeline fill / drain): nothing can overlap the upload of the very first shard or the download of the very last one,
    so those are kept small (`ramp` x the middle ones; equal shards when ramp >= 1 or there are too few shards)."""
    wts = [1.0] * nshards
    if nshards >= 3 * nworkers and 0 < ramp < 1:
        for k in range(nworkers):
            f = ramp + (1 - ramp) * k / nworkers
            wts[k] = f
            wts[nshards - 1 - k] = f
    cum = np.concatenate([[0.0], np.cumsum(wts)]) / sum(wts)
    bounds = [int(round(n * c)) for c in cum]
    bounds[0], bounds[-1] = 0, n
    return bounds


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_leg(P, tables, n_sample, steps, warmup, seed=0):
    """Time the CPU oracle (test infrastructure, used here only as the reported baseline) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import oracle_context, oracle_backend
    ctx = oracle_context()
    oracle_backend().dll.ora_set_num_threads(int(os.cpu_count() or 1))      # torchrun pins OMP_NUM_THREADS=1
    cores = int(oracle_backend().dll.ora_num_threads())
    mp, el, ph, po = make_world(P, ctx, tables, int(1.5 * n_sample) + 1024, n_sample + 1024, n_sample // 4 + 1024)
    el.upload(synth_electrons_numpy(P, n_sample, seed, 1))
    ctx.set_rng(seed, 0)
    psh = pusher(P)
    t = 0.0
    psteps, sub, el_t = 0, 0, 0.0
    times = []
    for it in range(warmup + steps):
        n0 = len(el)
        t0 = time.perf_counter()
        t += DT
        P.advance(mp, psh, t)
        for q in mp:
            P.droplow(q)
        dtm = time.perf_counter() - t0
        if it >= warmup:
            psteps += n0
            sub += P.last_advance_stats(mp)["substeps"]
            el_t += dtm
            times.append(dtm)
    ctx.close()
    return {"value": psteps / el_t, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_sample} electrons of the same RREA spectrum x {steps} steps (after {warmup} warm-up) through the C "
                      f"restatement of the reference algorithm (oracle/), OpenMP static schedule; Julia is not installable here",
            "kappa": sub / max(psteps, 1), "substeps_per_s": sub / el_t, "ms_per_step": 1e3 * el_t / max(len(times), 1)}


def _timed_steps(torch, P, mp, pops, psh, dt, steps, warmup, t0=0.0):
    """`steps` timed steps (advance! + droplow! per step) after `warmup`; returns device ms, particle-steps, sub-steps, stats."""
    t = t0
    ms = 0.0
    psteps = sub = 0
    kern_ms = kern_rows = 0.0
    per = []
    st = None
    for it in range(warmup + steps):
        n0 = sum(len(q) for q in pops)
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
            m = e0.elapsed_time(e1)
            ms += m; psteps += n0; sub += st["substeps"]; kern_ms += st["main_ms"]; kern_rows += st["main_rows"]; per.append(round(m, 3))
    return {"ms": ms, "psteps": psteps, "substeps": sub, "kern_ms": kern_ms, "kern_rows": kern_rows, "per_step_ms": per,
            "passes": st["passes"] if st else 0, "t": t}


def _probe_record(r, steps, peak, extra=None):
    out = {"particle_steps_per_s": r["psteps"] / (r["ms"] * 1e-3), "substeps_per_s": r["substeps"] / (r["ms"] * 1e-3),
           "kappa": r["substeps"] / max(r["psteps"], 1), "ms_per_step": r["ms"] / steps, "steps": steps, "passes_last_step": r["passes"],
           "hbm_frac_whole_step": ALGO_BYTES_PER_PARTICLE_STEP * r["psteps"] / (r["ms"] * 1e-3) / 1e9 / peak,
           "main_kernel_ms": r["kern_ms"] / steps,
           "hbm_frac_main_kernel": (ALGO_BYTES_PER_PARTICLE_STEP * r["kern_rows"] / (r["kern_ms"] * 1e-3) / 1e9 / peak) if r["kern_ms"] > 0 else None}
    out.update(extra or {})
    return out


def secondary_probes(torch, P, tables, peak, scale=1.0):
    """The other BASELINE.json configurations, a few device-timed steps each, every one with its HBM fraction
    (algorithmic 162 B per particle-step against the measured copy bandwidth) for the whole step and for the dominant kernel."""
    co = P.co
    out = {"note": "device-timed advance!+droplow! per step, populations resident in HBM, 3 timed steps after 2 warm-up steps; "
                   "hbm_frac_* = 162 B x particle-steps / time / measured HBM peak"}
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(11)

    def directions(n, cmin=-1.0):
        cost = rng.uniform(cmin, 1, n); phi = rng.uniform(0, 2 * np.pi, n); sint = np.sqrt(1 - cost ** 2)
        return np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)

    def state(sp, K, d):
        return dict(x=np.zeros((len(K), 3)), p=d * P.momentum_norm_from_kin(sp, K)[:, None], s=-np.log(1 - rng.random(len(K))))

    def world(ctx, ne, ng, npos, st_e=None, st_g=None, st_p=None, tabs=tables):
        el = P.Population(ctx, P.ELECTRON, int(1.8 * ne) + (1 << 18), st_e, tabs["electron"], 1e3 * co.eV)
        ph = P.Population(ctx, P.PHOTON, int(1.5 * ng) + (1 << 20), st_g, tabs["photon"], 1e3 * co.eV)
        po = P.Population(ctx, P.POSITRON, int(4 * npos) + (1 << 18), st_p, tabs["positron"], 1e2 * co.eV)
        return P.MultiPopulation(("electron", el), ("photon", ph), ("positron", po)), el, ph, po

    psh = pusher(P)
    # (1) photons only: the clean HBM-roofline measurement
    n = max(int(20_000_000 * scale), 4096)
    ctx = P.Context(device=torch.cuda.current_device(), stream=stream); ctx.set_profiling(True)
    Kg = np.exp(rng.uniform(np.log(1e4), np.log(3e7), n)) * co.eV
    mp, el, ph, po = world(ctx, 1 << 20, n, 1 << 16, st_g=state(P.PHOTON, Kg, directions(n)))
    # kernel figure: species of a pass launched one after the other (the default overlaps them on forked streams, and the
    # streaming kernel then shares the SMs with the collision chains of the secondary electrons: its event time is no longer
    # a bandwidth measurement); whole-step figure: the default, overlapped
    ctx.set_option("overlap", 0)
    rk = _timed_steps(torch, P, mp, (ph,), psh, DT, 3, 2)
    ctx.set_option("overlap", 1)
    r = _timed_steps(torch, P, mp, (ph,), psh, DT, 3, 1, t0=rk["t"])
    rec = _probe_record(r, 3, peak, {"workload": f"{n} photons, E ~ 1/E on [10 keV, 30 MeV], isotropic; their secondary electrons / positrons are "
                                                  "advanced in the same step (chains of up to ~400 collisions per electron bound the step)",
                                     "kernel": "k_advance_stream<photon>"})
    rec["main_kernel_ms"] = rk["kern_ms"] / 3
    rec["hbm_frac_main_kernel"] = (ALGO_BYTES_PER_PARTICLE_STEP * rk["kern_rows"] / (rk["kern_ms"] * 1e-3) / 1e9 / peak) if rk["kern_ms"] > 0 else None
    rec["ms_per_step_species_in_sequence"] = rk["ms"] / 3
    rec["kernel_timing"] = "streaming kernel timed with the species of a pass launched in sequence; ms_per_step with the default overlap"
    # the kernel does not rewrite the columns a free flight leaves unchanged (p, w, r): it moves 110.6 B per row (ncu,
    # profiles/r2_photon_stream_kernel_ncu_summary.csv), not the 162 algorithmic bytes, so the algorithmic fraction can exceed 1
    rec["dram_bytes_per_row_ncu"] = 110.6
    rec["hbm_frac_main_kernel_by_traffic"] = (110.6 * rk["kern_rows"] / (rk["kern_ms"] * 1e-3) / 1e9 / peak) if rk["kern_ms"] > 0 else None
    out["photon_streaming"] = rec
    ctx.close()
    # (2) electrons at kappa ~ 1 (dt scaled down): where HBM binds for leptons
    n = max(int(10_000_000 * scale), 4096)
    dt1 = DT / 2048
    ctx = P.Context(device=torch.cuda.current_device(), stream=stream); ctx.set_profiling(True)
    Fdt = co.elementary_charge * EFIELD * dt1
    tabs1 = dict(tables); tabs1["electron"] = P.build_electron_collision_table(P.air_composition(), Fdt, safety=1.15)
    mp, el, ph, po = world(ctx, 2 * n, n // 4, n // 16, tabs=tabs1)
    synth_electrons_device(torch, P, el, n, seed=7, uid0=1)
    r = _timed_steps(torch, P, mp, (el,), psh, dt1, 3, 3)
    out["electrons_kappa1"] = _probe_record(r, 3, peak, {"workload": f"{n} RREA electrons, dt = {dt1:.3e} s", "kernel": "k_advance_stream<electron>"})
    ctx.close()
    # (3) mixed feedback population, BASELINE configs[3]
    ne = ng = max(int(10_000_000 * scale), 4096); npos = max(int(200_000 * scale), 512)
    ctx = P.Context(device=torch.cuda.current_device(), stream=stream); ctx.set_profiling(True)
    Ke = np.clip(rng.exponential(7.3e6, ne), 1e3 * 1.0001, 1e8) * co.eV
    Kg = np.exp(rng.uniform(np.log(1e4), np.log(3e7), ng)) * co.eV
    Kp = np.exp(rng.uniform(np.log(1e5), np.log(2e7), npos)) * co.eV
    mp, el, ph, po = world(ctx, ne, ng, npos, state(P.ELECTRON, Ke, directions(ne, 0.8)), state(P.PHOTON, Kg, directions(ng)),
                           state(P.POSITRON, Kp, directions(npos)))
    r = _timed_steps(torch, P, mp, (el, ph, po), psh, DT, 3, 2)
    out["mixed_config4"] = _probe_record(r, 3, peak, {"workload": f"{ne} e- + {ng} gamma + {npos} e+, all nine processes (BASELINE configs[3])",
                                                      "flags": int(ctx.error_flags())})
    ctx.close()
    # (4) LXCat slow-electron swarm, 64 channels, BASELINE configs[4]
    n = max(int(10_000_000 * scale), 4096)
    ctx = P.Context(device=torch.cuda.current_device(), stream=stream); ctx.set_profiling(True)
    tab = P.synthetic_lxcat_table(grid_kind=0, extra_levels=57)
    stt = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * co.eV / co.electron_mass) * 1.2, s=-np.log(1 - rng.random(n)))
    pop = P.Population(ctx, P.SLOW_ELECTRON, int(1.5 * n), stt, tab, 0.0)
    mps = P.MultiPopulation(("slow", pop))
    pshs = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -100 * co.Td * co.nair]), None))
    r = _timed_steps(torch, P, mps, (pop,), pshs, 1e-12, 3, 2)
    out["lxcat64_config5"] = _probe_record(r, 3, peak, {"workload": f"{n} slow electrons, Maxwellian 2 eV, 100 Td, dt = 1e-12 s, {len(tab.proc)} channels on a linear table",
                                                        "kernel": "k_advance_wq<slow_electron, linear>"})
    ctx.close()
    # (5) latency regime: the reference's own working size (scripts/swarm.jl:29-30 keeps 1e4 electrons)
    n = 10_000
    ctx = P.Context(device=torch.cuda.current_device(), stream=stream); ctx.set_profiling(True)
    mp, el, ph, po = world(ctx, 1 << 18, 1 << 18, 1 << 16)
    synth_electrons_device(torch, P, el, n, seed=3, uid0=1)
    r = _timed_steps(torch, P, mp, (el,), psh, DT, 5, 3)
    out["swarm_1e4_latency"] = _probe_record(r, 5, peak, {"workload": "1e4 RREA electrons (the reference keeps its swarm at 1e4 electrons by roulette)",
                                                          "per_step_ms": r["per_step_ms"]})
    ctx.close()
    return out


def strong_leg(torch, dist, P, ctx, mp, el, psh, args, rank, world, barrier):
    """BASELINE configs[2] as named: 1e8 electrons in TOTAL, split over the N GPUs (slightly unevenly, so that the rebalance
    moves rows), with the library's NCCL collectives inside the timed region."""
    from particulator_b200 import dist as pdist
    total = args.strong_total
    share = [1.0 + 0.2 * ((r / (world - 1)) - 0.5) for r in range(world)]          # +-10 % around the mean
    counts = [int(total * sh / sum(share)) for sh in share]
    n = counts[rank]
    synth_electrons_device(torch, P, el, n, seed=4321 + rank, uid0=1 + rank * (1 << 40))
    t = 100.0 * DT            # any start time: t is a per-particle column, set below
    column_tensor(torch, el, 7, n).fill_(t)
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    coll_ms = 0.0
    diag_ms = 0.0
    wait_ms = 0.0
    moved_rows = 0
    psteps = 0

    def one(it, timed):
        nonlocal t, coll_ms, diag_ms, wait_ms, moved_rows, psteps
        n0 = len(el)
        t += DT
        P.advance(mp, psh, t)
        for q in mp:
            P.droplow(q)
        # what a rank spends from the end of its own step to the end of the collectives has two parts: WAITING for the slowest
        # rank (load imbalance: it would wait at the next synchronisation point anyway) and the collectives themselves.  A
        # barrier in front separates the two.
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        dist.barrier()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        d = pdist.diag_allreduce(el)                      # global counts / moments every step (run.jl:31-40)
        w2 = time.perf_counter()
        if it % 2 == 1:
            _, mv = pdist.rebalance_device(el, tolerance=0.02)
            if timed:
                moved_rows += abs(mv)
        ctx.synchronize()
        w3 = time.perf_counter()
        if timed:
            wait_ms += (w1 - w0) * 1e3
            diag_ms += (w2 - w1) * 1e3
            coll_ms += (w3 - w1) * 1e3
            psteps += n0
        return d

    for it in range(warmup):
        one(it, False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d = None
    for it in range(steps):
        d = one(warmup + it, True)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1), coll_ms, diag_ms, wait_ms], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(psteps), float(moved_rows), float(len(el))], device="cuda", dtype=torch.float64)
    nmax = torch.tensor([float(len(el))], device="cuda", dtype=torch.float64)
    nmin = nmax.clone()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
    return {"scaling": "strong", "electrons_total": total, "electrons_per_gpu_start": counts, "steps": steps,
            "value": float(tot[0]) / (float(ms[0]) * 1e-3), "unit": "particle-steps/s", "ms_per_step": float(ms[0]) / steps,
            "collective_ms_per_step": float(ms[1]) / steps, "diag_allreduce_ms_per_step": float(ms[2]) / steps,
            "rebalance_ms_per_step": (float(ms[1]) - float(ms[2])) / steps, "wait_for_slowest_rank_ms_per_step": float(ms[3]) / steps,
            "rebalanced_rows_per_step": float(tot[1]) / 2 / steps,
            "global_n_from_diag_allreduce": int(d.n), "global_n_from_sum": int(tot[2]),
            "n_per_gpu_end": {"max": int(nmax.item()), "min": int(nmin.item())},
            "collectives": "ptl_diag_allreduce every step + ptl_rebalance (tolerance 2 %) every second step, NCCL called from the library, inside the timed "
                           "region; *_ms_per_step are host wall times, max over ranks; a barrier in front of the collectives separates waiting for the slowest "
                           "rank (load imbalance) from the collectives themselves"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=int(os.environ.get("PTL_BENCH_N", 100_000_000)))
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--strong-total", type=int, default=int(os.environ.get("PTL_BENCH_STRONG_TOTAL", 100_000_000)))
    ap.add_argument("--secondary-scale", type=float, default=1.0, help="scale the secondary probe populations (tests use < 1)")
    ap.add_argument("--e2e-shards", type=int, default=int(os.environ.get("PTL_E2E_SHARDS", 12)))
    ap.add_argument("--e2e-advance-slots", type=int, default=int(os.environ.get("PTL_E2E_ADVANCE_SLOTS", 2)),
                    help="how many workers of the e2e leg may be inside advance!+droplow! at the same time (0 = all of them)")
    ap.add_argument("--e2e-ramp", type=float, default=float(os.environ.get("PTL_E2E_RAMP", 0.3)),
                    help="relative size of the first and last shard of the e2e leg (1 = equal shards): small shards at both ends shorten "
                         "the time before the first kernel can start and the last download after the last kernel")
    ap.add_argument("--e2e-workers", type=int, default=int(os.environ.get("PTL_E2E_WORKERS", 4)))
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    import particulator_b200 as P
    config = {"workload": "RREA electron avalanche in STP air (BASELINE.json configs[2]): E~exp(-E/7.3MeV) on [1 keV,100 MeV], "
                          "uniform E_z=-5e5 V/m, dt=2.5e-11 s, e-/gamma/e+ populations, advance!+droplow! per step",
              "electrons_per_gpu": args.n_per_gpu, "electrons_total": args.n_per_gpu * max(world, 1), "dt_s": DT,
              "parallelism": f"particle-sharded x{max(world, 1)}, no data-path collective",
              "l2_policy": "inputs larger than L2 (>= 8.9 GB of particle state per GPU)"}

    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        tables = build_tables(P)
        res = cpu_leg(P, tables, args.cpu_sample, max(args.steps, 1), max(args.warmup, 1))
        config["electrons_per_gpu"] = args.cpu_sample
        config["electrons_total"] = args.cpu_sample
        config["note"] = "reference arm = CPU restatement of the reference algorithm on a bounded sample of the same workload"
        line = {"metric": "particle-steps/sec", "value": res["value"], "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "kappa": res["kappa"], "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------------
    # exactly one JSON line on stdout: everything else this process (or NCCL) prints goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tables = build_tables(P)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = P.Context(device=local_rank, stream=stream)
    ctx.set_profiling(True)
    n = args.n_per_gpu
    cap_e = int(1.6 * n) + 4096       # room for the in-step births (~10 %) and the quasi-steady keV secondaries
    mp, el, ph, po = make_world(P, ctx, tables, cap_e, max(n // 2, 1 << 20), max(n // 16, 1 << 18))
    uid0 = 1 + rank * (1 << 40)
    synth_electrons_device(torch, P, el, n, seed=1234 + rank, uid0=uid0)      # (the synthetic INPUT differs per rank; the RNG seed does not)
    ctx.set_rng(0, 0)              # ONE seed for the whole job: Philox streams are keyed by uid, so shards are independent through
                                   # their disjoint uid ranges and the result does not depend on how many GPUs share the particles
    if dist is not None:
        from particulator_b200 import dist as pdist
        pdist.init_comm(ctx, dist)                                            # NCCL communicator inside the library (ptl_comm_init)
    psh = pusher(P)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    t = 0.0

    host_ms = {"advance": [], "droplow": []}     # host wall time of the two calls (both end in a scalar read-back)

    def step():
        nonlocal t
        n0 = len(el)
        t += DT
        h0 = time.perf_counter()
        P.advance(mp, psh, t)
        h1 = time.perf_counter()
        st = P.last_advance_stats(mp)
        for q in mp:
            P.droplow(q)
        h2 = time.perf_counter()
        host_ms["advance"].append((h1 - h0) * 1e3)
        host_ms["droplow"].append((h2 - h1) * 1e3)
        return n0, st

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # started before the warm-up so that its NVML initialisation is over when the timing starts
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.t_begin = time.time()  # only samples taken inside the timed region are reported
    ctx.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    psteps = substeps = 0
    main_ms = main_rows = 0.0
    host_ms["advance"].clear(); host_ms["droplow"].clear()
    for _ in range(args.steps):
        n0, st = step()
        psteps += n0
        substeps += st["substeps"]
        main_ms += st["main_ms"]
        main_rows += st["main_rows"]
    e1.record()
    barrier()
    host_steps = {k: [round(v, 1) for v in vals] for k, vals in host_ms.items()}
    elapsed_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count()
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([float(psteps), float(substeps), float(launches)], device="cuda", dtype=torch.float64)
    mx = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    psteps_all, substeps_all, launches_all = (float(v) for v in tot.tolist())
    elapsed_all = float(mx.item())
    value = psteps_all / (elapsed_all * 1e-3)

    # ---- roofline of the dominant kernel (this rank) ----
    peak, peak_src = measured_hbm_peak()
    achieved = (ALGO_BYTES_PER_PARTICLE_STEP * main_rows) / (main_ms * 1e-3) / 1e9 if main_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_row"] * (main_rows / max(args.steps, 1))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": peak_src, "kernel": "k_advance_wq<electron> (first pass)",
                "kernel_ms_per_launch": main_ms / max(args.steps, 1), "rows_per_launch": main_rows / max(args.steps, 1),
                "algorithmic_bytes_per_row": ALGO_BYTES_PER_PARTICLE_STEP,
                "note": "electrons in STP air do kappa collision sub-steps per particle-step; the kernel is instruction-issue bound, "
                        "not HBM bound, whenever kappa >> 1 (see DESIGN.md section 6); kappa is reported below"}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        import psutil
        n_now = len(el)
        avail = psutil.virtual_memory().available / max(world, 1) * 0.8   # every rank of the box pins its own copies at the same time
        if avail < 3.2 * 89 * n_now:                                  # host RAM guard: e2e keeps two pinned copies
            keep = int(avail / (3.2 * 89))
            el.set_n(keep)
            config["e2e_note"] = f"host RAM limited the e2e leg to {keep} of {n_now} electrons"
        d = el.download()
        n_e2e = len(d["w"])
        host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in d.items()}
        del d
        import ctypes as C
        from particulator_b200._lib import AdvanceStats
        b = ctx.backend

        def ptr(tn, ty, off=0):
            return C.cast(tn.data_ptr() + off * tn.element_size() * (tn.shape[1] if tn.dim() > 1 else 1), C.POINTER(ty))

        # The population is cut into independent shards (particles never interact) that two host threads push through
        # their own library contexts: while one context advances a shard, the other one's H2D / D2H copies run on the
        # copy engines.  Every byte still crosses PCIe inside the timed region.
        nshards, nworkers = max(args.e2e_shards, 1), max(args.e2e_workers, 1)
        bounds = shard_bounds(n_e2e, nshards, nworkers, args.e2e_ramp)
        shard_cap = int(1.7 * max(bounds[k + 1] - bounds[k] for k in range(nshards))) + 8192
        workers = []
        for wk in range(nworkers):
            wctx = P.Context(device=local_rank)                       # own non-blocking stream
            wmp, wel, wph, wpo = make_world(P, wctx, tables, shard_cap, max(shard_cap // 3, 1 << 20), max(shard_cap // 16, 1 << 18))
            wctx.set_rng(0, 0)
            wctx.set_profiling(True)
            workers.append((wctx, wmp, wel, psh.desc(wctx)))
        got_rows = [0] * nshards
        out_cap = [int(1.3 * (bounds[k + 1] - bounds[k])) + 4096 for k in range(nshards)]
        out = [{k: torch.empty((out_cap[sh],) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in host.items()} for sh in range(nshards)]
        t_loc = float(host["t"][0]) + DT
        errors = []

        phase_ms = [[0.0, 0.0, 0.0] for _ in range(nworkers)]      # upload / advance+droplow / download wall time per worker (last step)
        timeline = []                                               # (worker, shard, t_upload, t_advance, t_download, t_end) of the last step, ms
        step_t0 = [0.0]

        # The main kernel of a shard fills every SM, so the kernels of different workers run one after the other whatever the
        # host does.  Unthrottled, the short newborn passes of worker A queue behind the main kernels B and C launched in the
        # meantime, all workers end their advance! together and then copy together while the GPU idles (shard timeline in the
        # record).  A semaphore lets only `slots` workers into advance! at a time; the others use the wait for their copies.
        slots = args.e2e_advance_slots if args.e2e_advance_slots > 0 else nworkers
        adv_sem = threading.Semaphore(slots)

        def work(wk):
            wctx, wmp, wel, pd = workers[wk]
            phase_ms[wk][:] = [0.0, 0.0, 0.0]
            try:
                for sh in range(wk, nshards, nworkers):
                    lo, m = bounds[sh], bounds[sh + 1] - bounds[sh]
                    if m <= 0:                                         # tiny runs: a ramped shard may be empty
                        continue
                    tp0 = time.perf_counter()
                    rc = b.population_upload(wctx.h, wel.id, m, ptr(host["x"], C.c_double, lo), ptr(host["p"], C.c_double, lo),
                                             ptr(host["w"], C.c_double, lo), ptr(host["t"], C.c_double, lo), ptr(host["s"], C.c_double, lo),
                                             ptr(host["r"], C.c_double, lo), ptr(host["active"], C.c_uint8, lo), ptr(host["uid"], C.c_uint64, lo))
                    assert rc == 0, rc
                    with adv_sem:
                        tp1 = time.perf_counter()
                        rc = b.advance(wctx.h, wmp.id, C.byref(pd), t_loc, None)
                        assert rc >= 0, rc
                        for q in wmp:
                            b.droplow(wctx.h, q.id, 0.0)
                        wctx.synchronize()
                        tp2 = time.perf_counter()
                    ast_ = AdvanceStats()
                    b.last_advance_stats(wctx.h, C.byref(ast_))
                    o = out[sh]
                    got_rows[sh] = int(b.population_download(wctx.h, wel.id, out_cap[sh], ptr(o["x"], C.c_double), ptr(o["p"], C.c_double),
                                                             ptr(o["w"], C.c_double), ptr(o["t"], C.c_double), ptr(o["s"], C.c_double),
                                                             ptr(o["r"], C.c_double), ptr(o["active"], C.c_uint8), ptr(o["uid"], C.c_uint64)))
                    assert got_rows[sh] > 0
                    tp3 = time.perf_counter()
                    phase_ms[wk][0] += (tp1 - tp0) * 1e3; phase_ms[wk][1] += (tp2 - tp1) * 1e3; phase_ms[wk][2] += (tp3 - tp2) * 1e3
                    timeline.append((wk, sh, round((tp0 - step_t0[0]) * 1e3, 1), round((tp1 - step_t0[0]) * 1e3, 1),
                                     round((tp2 - step_t0[0]) * 1e3, 1), round((tp3 - step_t0[0]) * 1e3, 1),
                                     round(ast_.main_ms, 1), int(ast_.passes)))
                    # photons / positrons born in this shard stay on the device (they are results of later steps' inputs)
                    b.population_clear(wctx.h, list(wmp)[1].id)
                    b.population_clear(wctx.h, list(wmp)[2].id)
            except Exception as exc:  # pragma: no cover
                errors.append(repr(exc))

        def one_e2e_step():
            del timeline[:]
            step_t0[0] = time.perf_counter()
            ths = [threading.Thread(target=work, args=(wk,)) for wk in range(nworkers)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()

        one_e2e_step()                                                # untimed warm-up (lazy module loading, allocations)
        barrier()
        t0 = time.perf_counter()
        e2e_psteps = 0
        d2h_rows = 0
        for k in range(args.e2e_steps):
            one_e2e_step()
            e2e_psteps += n_e2e
            d2h_rows += sum(got_rows)
        torch.cuda.synchronize()
        e2e_elapsed = (time.perf_counter() - t0) * 1e3
        barrier()
        assert not errors, errors
        e2e_ms = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        e2e_ps = torch.tensor([float(e2e_psteps)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(e2e_ps, op=dist.ReduceOp.SUM)
        e2e = {"value": float(e2e_ps.item()) / (float(e2e_ms.item()) * 1e-3), "unit": "particle-steps/s",
               "h2d_bytes_per_step": 89 * n_e2e, "d2h_bytes_per_step": 89 * d2h_rows // max(args.e2e_steps, 1), "steps": args.e2e_steps,
               "ms_per_step": float(e2e_ms.item()) / max(args.e2e_steps, 1),
               "path": "pinned host arrays -> ptl_population_upload -> ptl_advance -> ptl_droplow -> ptl_population_download, "
                       f"{nshards} independent shards through {nworkers} contexts (copies of one overlap the kernels of the other; first/last shards {args.e2e_ramp:g}x the middle ones; {slots} worker(s) inside advance! at a time)",
               "timer": "host wall clock between device-wide synchronizes (work spans several streams)",
               "worker_phase_ms_last_step": {"upload": [round(p[0], 1) for p in phase_ms], "advance_droplow": [round(p[1], 1) for p in phase_ms],
                                             "download": [round(p[2], 1) for p in phase_ms]},
               "shard_timeline_ms_last_step": {"columns": ["worker", "shard", "upload_start", "advance_start", "download_start", "end", "main_kernel_ms", "passes"],
                                               "rows": sorted(timeline, key=lambda r: r[2])}}
        for wctx, *_ in workers:
            wctx.close()

    # ---- CPU baseline on the host cores of this box (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_leg(P, tables, args.cpu_sample, max(args.cpu_steps, 1), 1)

    # ---- secondary: the other BASELINE configurations and the latency regime (rank 0, N = 1 only) ----
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        try:
            secondary = secondary_probes(torch, P, tables, peak, args.secondary_scale)
        except Exception as exc:  # pragma: no cover  (a failed probe must not void the headline)
            secondary = {"error": repr(exc)}

    # ---- strong scaling with the collectives inside the timed region (N > 1 only) ----
    strong = None
    if world > 1 and not args.no_strong:
        for q in mp:
            q.set_n(0)
        strong = strong_leg(torch, dist, P, ctx, mp, el, psh, args, rank, world, barrier)

    if rank == 0:
        line = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": max(world, 1), "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_all / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches_all), "roofline": roofline, "cpu_baseline": cpu,
                "kappa": substeps_all / max(psteps_all, 1.0), "substeps_per_s": substeps_all / (elapsed_all * 1e-3),
                "hbm_roofline_frac_whole_step": (ALGO_BYTES_PER_PARTICLE_STEP * value / max(world, 1)) / 1e9 / peak,
                "host_ms_per_step": host_steps, "secondary": secondary, "strong": strong}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
