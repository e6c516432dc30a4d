"""2+ GPU validation of the multi-GPU entry points of the C ABI (run under torchrun): shards of very different sizes are
advanced, diagnostics are all-reduced and the electron population is rebalanced by the LIBRARY's NCCL calls
(ptl_comm_init, ptl_diag_allreduce, ptl_rebalance; torch.distributed only ships the 128-byte id and gathers the uids for the
check), and the result is checked: every uid exists exactly once, global weight is conserved,
shard sizes are equal, and a further advance! works on the rebalanced shards."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import particulator_b200 as P
from particulator_b200 import dist as pdist
import bench


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    tabs = bench.build_tables(P)
    ctx = P.Context(device=lr, stream=torch.cuda.current_stream().cuda_stream)
    n = 400_000 * (1 + 3 * rank)                      # unbalanced on purpose
    mp, el, ph, po = bench.make_world(P, ctx, tabs, 4_000_000, 1 << 20, 1 << 18)
    bench.synth_electrons_device(torch, P, el, n, seed=10 + rank, uid0=1 + rank * (1 << 40))
    ctx.set_rng(7, 0)                                  # one seed for every rank: streams are keyed by uid, never by rank
    pdist.init_comm(ctx, dist)
    psh = bench.pusher(P)
    P.advance(mp, psh, bench.DT)
    for q in mp:
        P.droplow(q)
    d = pdist.diag_allreduce(el)                       # global sums through the library's communicator
    vs = torch.tensor([float(d.nactive), d.weight, d.wenergy], device="cuda", dtype=torch.float64)
    n_before = len(el)
    n_after, moved = pdist.rebalance_device(el, tolerance=0.02)
    uids = torch.as_tensor(pdist._DevArray(el.column_ptr(11), n_after, "<i8"), device="cuda").clone()
    sizes = pdist.gather_counts(dist, n_after, device="cuda")
    mx = max(sizes)
    padded = torch.full((mx,), -1, dtype=torch.int64, device="cuda")
    padded[:n_after] = uids
    gathered = [torch.zeros(mx, dtype=torch.int64, device="cuda") for _ in sizes]
    dist.all_gather(gathered, padded)
    d2 = pdist.diag_allreduce(el)
    vs2 = torch.tensor([float(d2.nactive), d2.weight, d2.wenergy], device="cuda", dtype=torch.float64)
    allu = torch.cat([g[:k] for g, k in zip(gathered, sizes)])
    ok_unique = bool(len(torch.unique(allu)) == len(allu) == sum(sizes))
    P.advance(mp, psh, 2 * bench.DT)                   # the rebalanced shards keep working
    st = P.last_advance_stats(mp)
    flags = ctx.error_flags()
    if rank == 0:
        print(json.dumps({"world": world, "sizes_after": sizes, "n_before_rank0": n_before, "n_after_rank0": n_after,
                          "global_before": vs.tolist(), "global_after": vs2.tolist(), "uids_unique": ok_unique,
                          "conserved": bool(abs(vs2[1].item() - vs[1].item()) < 1e-9 * vs[1].item() and abs(vs2[2].item() - vs[2].item()) < 1e-9 * abs(vs[2].item())),
                          "balanced": max(sizes) - min(sizes) <= 1, "substeps_after": st["substeps"], "flags": flags}))
    pdist.destroy_comm(ctx)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
