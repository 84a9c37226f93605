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
    """Slice every per-problem array of a scenario batch (obca_b200.scenarios: parking or quadcopter) for this rank: every
    numpy array whose leading dimension is the batch size B."""
    B = sc["B"]
    lo, hi = shard_range(B, world, rank)
    out = dict(sc)
    for k, v in sc.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B and k not in ("A", "b", "vOb", "ego", "XYbounds", "obs"):
            out[k] = v[lo:hi]
    out["B"] = hi - lo
    out["offset"] = lo
    return out


def reduce_stats(dist, device, sums: dict, maxes: dict) -> dict:
    """The collectives of the path: ONE all-reduce (SUM) of the counters and ONE (MAX) of the times.  dist = torch.distributed
    (nccl on GPUs, gloo in the CPU tests) or None for a single process."""
    import torch
    ks, km = list(sums), list(maxes)
    ts = torch.tensor([float(sums[k]) for k in ks], dtype=torch.float64, device=device)
    tm = torch.tensor([float(maxes[k]) for k in km], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out = {k: float(v) for k, v in zip(ks, ts.tolist())}
    out.update({k: float(v) for k, v in zip(km, tm.tolist())})
    return out


def reduce_counters(dist, device, converged: int, iterations: int, problems: int, seconds: float):
    """converged / iterations / problems summed over the ranks, device time = max over the ranks, whole-job trajectories per second."""
    r = reduce_stats(dist, device, dict(converged=converged, iterations=iterations, problems=problems), dict(seconds=seconds))
    return dict(converged=int(r["converged"]), iterations=int(r["iterations"]), problems=int(r["problems"]), seconds=r["seconds"],
                traj_per_s=r["converged"] / max(r["seconds"], 1e-300))


def pin_rank_to_cpus(local_rank: int, local_world: int):
    """One disjoint slice of the host's CPUs per rank (the solver's host loop issues a few hundred driver calls per step; ranks that
    share cores pace each other).  Returns the CPU list or None when the platform does not support affinities."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(local_world, 1)
        if per < 1:
            return None
        mine = cpus[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None
