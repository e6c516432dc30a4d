"""Checkpoint / restart of a MultiPopulation ("checkpoint = SoA dump + step + seed").

The reference has no checkpoint code (HDF5/JLD2 appear only as unused script dependencies); what a restart needs follows
from what `advance!` reads: every column of every population (`x, p, w, t, s, r, active`, population.jl:7-21), the
simulation time, and — because the RNG here is counter-based — the (seed, advance-call index) pair plus the per-particle
`uid` column that keys each Philox stream.  With those restored, the continued run is the same run: the test replays
k steps + checkpoint + k steps against 2k uninterrupted steps and requires identical state.

File format (one file, little-endian): numpy `.npz` container with
    meta                 JSON bytes: {"format": "particulator_b200.checkpoint", "version": 1, "t": ..., "seed": ..., "step": ..., "next_uid": ...,
                          "populations": [{"name", "species", "n", "capacity", "energy_cut"}, ...]}
    <name>/x, <name>/p   float64 [n,3], xyz-interleaved like the reference's Vector{SVector{3,Float64}}
    <name>/w,t,s,r       float64 [n];   <name>/active uint8 [n];   <name>/uid uint64 [n]
Rows are saved as they lie in the store (inactive rows included), so row order, `n` and the compaction state survive.
Collision tables and pushers are not saved: they are rebuilt by the script, as in the reference.
"""
import io
import json

import numpy as np

FORMAT = "particulator_b200.checkpoint"
VERSION = 1
_COLUMNS = ("x", "p", "w", "t", "s", "r", "active", "uid")


def save_checkpoint(path, mpopl, t, extra=None):
    """Write every population of `mpopl` plus the RNG position of its context.  `path` may be a file name or a
    file-like object.  Returns the meta dict."""
    pops = list(mpopl.pairs())
    ctx = pops[0][1].ctx
    seed, step = ctx.get_rng()
    # next_uid: the counter behind default uids.  Without it a restored run would hand out uids that are still alive
    # (uids key the RNG streams: two particles with one uid draw the same collision deviates).
    meta = {"format": FORMAT, "version": VERSION, "t": float(t), "seed": int(seed), "step": int(step),
            "next_uid": int(ctx.get_uid_counter()), "populations": [], "extra": extra or {}}
    arrays = {}
    for name, popl in pops:
        d = popl.download()
        n = len(d["w"])
        meta["populations"].append({"name": str(name), "species": int(popl.species), "n": int(n),
                                    "capacity": int(popl.capacity), "energy_cut": float(popl.energy_cut)})
        for c in _COLUMNS:
            arrays[f"{name}/{c}"] = d[c]
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__"):
        with open(path, "wb") as f:          # np.savez(name) would append ".npz" to a name without it
            np.savez(f, **arrays)
    else:
        np.savez(path, **arrays)
    return meta


def read_checkpoint(path):
    """Parse a checkpoint file: returns (meta, {name: {column: array}})."""
    with np.load(path) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format") != FORMAT or meta.get("version") != VERSION:
            raise ValueError(f"not a {FORMAT} v{VERSION} file")
        state = {p["name"]: {c: z[f"{p['name']}/{c}"] for c in _COLUMNS} for p in meta["populations"]}
    for p in meta["populations"]:
        st = state[p["name"]]
        if any(len(st[c]) != p["n"] for c in _COLUMNS):
            raise ValueError(f"population {p['name']}: column length differs from n = {p['n']}")
    return meta, state


def load_checkpoint(path, mpopl):
    """Restore `mpopl` (built by the script with the same tables) from a checkpoint: contents of every population, and the
    RNG position of the context.  Returns the saved time `t`."""
    meta, state = read_checkpoint(path)
    pops = dict((str(k), v) for k, v in mpopl.pairs())
    missing = sorted(set(pops) - {p["name"] for p in meta["populations"]})
    if missing:
        raise KeyError(f"the MultiPopulation has populations the checkpoint does not hold: {missing}")
    for p in meta["populations"]:
        if p["name"] not in pops:
            raise KeyError(f"checkpoint has population {p['name']!r}, the MultiPopulation does not")
        popl = pops[p["name"]]
        if int(popl.species) != p["species"]:
            raise ValueError(f"population {p['name']}: species differs from the checkpoint")
        if p["n"] > popl.capacity:
            raise ValueError(f"population {p['name']}: {p['n']} saved rows exceed the capacity {popl.capacity}")
        if p["n"] > 0:
            popl.upload(state[p["name"]])
        else:
            popl.set_n(0)
    ctx = next(iter(pops.values())).ctx
    ctx.set_rng(meta["seed"], meta["step"])
    if "next_uid" in meta:
        ctx.set_uid_counter(max(int(meta["next_uid"]), ctx.get_uid_counter()))
    return meta["t"]


def dumps(mpopl, t, extra=None):
    """Checkpoint as bytes (for shipping between ranks or into an object store)."""
    buf = io.BytesIO()
    save_checkpoint(buf, mpopl, t, extra)
    return buf.getvalue()


def loads(data, mpopl):
    return load_checkpoint(io.BytesIO(data), mpopl)
