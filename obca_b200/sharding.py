"""Batch sharding across GPUs: the batch of independent trajectory problems is the only
shard axis.  One process per GPU; rank r owns the contiguous slice shard_range(B, world, r); nothing is exchanged
on the solve path; one all-reduce (SUM) carries {converged, iterations, problems} and one (MAX) the device time."""
from __future__ import annotations

import numpy as np


def shard_range(B: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition of range(B): the first B % world ranks get one extra problem."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(sc: dict, world: int, rank: int) -> dict:
    """Slice every per-problem array of a scenario batch (obca_b200.scenarios) for this rank."""
    lo, hi = shard_range(sc["B"], world, rank)
    out = dict(sc)
    for k in ("x0", "rx", "ry", "ryaw", "xWS", "uWS"):
        out[k] = sc[k][lo:hi]
    if np.ndim(sc["xF"]) == 2:
        out["xF"] = sc["xF"][lo:hi]
    out["B"] = hi - lo
    out["offset"] = lo
    return out


def reduce_counters(dist, device, converged: int, iterations: int, problems: int, seconds: float):
    """The single collective of the path.  dist = torch.distributed (nccl on GPUs, gloo in the CPU tests)."""
    import torch
    cnt = torch.tensor([converged, iterations, problems], dtype=torch.float64, device=device)
    tmax = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    c, i, p = [float(x) for x in cnt.tolist()]
    return dict(converged=int(c), iterations=int(i), problems=int(p), seconds=float(tmax.item()),
                traj_per_s=c / max(float(tmax.item()), 1e-300))
