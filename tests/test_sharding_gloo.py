"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a batch, "solve" their slice with a stand-in
for the device call, and all-reduce the counters exactly as bench.py / obca_b200.sharding do on NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from obca_b200 import scenarios, sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenarios.reverse_parking_batch(11, 20, seed=0)          # same global batch on every rank
    mine = sharding.shard_batch(sc, world, rank)
    # stand-in for the device solve: "converges" when X0 < 5 and takes (index + 1) iterations
    conv = int((mine["x0"][:, 0] < 5).sum()); iters = int(sum(mine["offset"] + i + 1 for i in range(mine["B"])))
    r = sharding.reduce_counters(dist, "cpu", conv, iters, mine["B"], 0.5 + rank)
    q.put((rank, mine["offset"], mine["B"], r))
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    for B in (1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_counters():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    sc = scenarios.reverse_parking_batch(11, 20, seed=0)
    assert [(r[1], r[2]) for r in res] == [(0, 6), (6, 5)]
    for _, _, _, r in res:                                         # every rank sees the same reduced values
        assert r["problems"] == 11 and r["converged"] == int((sc["x0"][:, 0] < 5).sum())
        assert r["iterations"] == sum(range(1, 12)) and r["seconds"] == 1.5
        assert abs(r["traj_per_s"] - r["converged"] / 1.5) < 1e-12


def test_shard_batch_slices_every_per_problem_array_and_pins_cpus():
    """Parking and quadcopter batches: every array with leading dimension B is sliced, shared data (A, b, obstacles) is not."""
    scq = scenarios.quadcopter_batch(7, 10, seed=2)
    a, b = sharding.shard_batch(scq, 2, 0), sharding.shard_batch(scq, 2, 1)
    assert a["B"] + b["B"] == 7 and a["x0"].shape == (4, 12) and b["xWS"].shape[0] == 3 and a["obs"].shape == scq["obs"].shape
    assert np.array_equal(np.concatenate([a["xF"], b["xF"]]), scq["xF"])
    scp = scenarios.reverse_parking_batch(5, 20, seed=0)
    c = sharding.shard_batch(scp, 5, 3)
    assert c["B"] == 1 and c["offset"] == 3 and np.array_equal(c["rx"][0], scp["rx"][3]) and c["A"].shape == scp["A"].shape
    r = sharding.reduce_stats(None, "cpu", dict(a=2, b=3), dict(t=0.25))
    assert r == dict(a=2.0, b=3.0, t=0.25)
    before = sorted(os.sched_getaffinity(0))
    try:
        mine = sharding.pin_rank_to_cpus(1, 2)
        assert mine is None or (len(mine) == len(before) // 2 and set(mine) <= set(before))
    finally:
        os.sched_setaffinity(0, before)
