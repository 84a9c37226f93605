"""The reference's warm-start pipeline around Hybrid A* (AutonomousParking/main.jl:99-108,111-205,215-248): obstacle point
clouds of the two demo scenarios, path -> speed profile -> smoothed speed / acceleration / steering -> down-sampling to the
NLP grid.  Output = exactly what main.jl hands to ParkingDist / ParkingSignedDist: (rx, ry, ryaw) sampled, xWS, uWS, N."""
from __future__ import annotations

import math
import os

import numpy as np

from . import hybrid_a_star
from .velo_smooth import velo_smooth


def _frange(a, b, step):
    """Julia's a:step:b for the 0.1-step ranges of main.jl (end point included when it is hit up to round-off)."""
    n = int(math.floor((b - a) / step + 1e-9))
    return [a + i * step for i in range(n + 1)]


def obstacle_points(scenario: str):
    """main.jl:111-142 ("backwards") / :170-205 ("parallel"): the point obstacles Hybrid A* plans against."""
    ox, oy = [], []
    if scenario == "backwards":
        for v in _frange(-12.0, -1.3, 0.1):
            ox.append(v); oy.append(5.0)
        for i in range(-2, 6):
            ox.append(-1.3); oy.append(float(i))
        for i in range(-2, 6):
            ox.append(1.3); oy.append(float(i))
        for v in _frange(1.3, 12.0, 0.1):
            ox.append(v); oy.append(5.0)
        for i in range(-12, 13):
            ox.append(float(i)); oy.append(11.0)
    elif scenario == "parallel":
        for v in _frange(-12.0, -3.0, 0.1):
            ox.append(v); oy.append(5.0)
        for i in range(-2, 6):
            ox.append(-3.0); oy.append(float(i))
        for i in range(-3, 4):
            ox.append(float(i)); oy.append(2.5)
        for i in range(-2, 6):
            ox.append(3.0); oy.append(float(i))
        for v in _frange(3.0, 12.0, 0.1):
            ox.append(v); oy.append(5.0)
        for i in range(-12, 13):
            ox.append(float(i)); oy.append(11.5)
    else:
        raise ValueError("scenario must be 'backwards' or 'parallel'")
    return np.array(ox), np.array(oy)


def warm_start_from_path(rx, ry, ryaw, Ts, L=2.7, sampleN=3, motionStep=0.1, amax=0.3):
    """main.jl:222-248.  (rx, ry, ryaw): Hybrid A* path, one point per motionStep.  Returns dict(rx, ry, ryaw, xWS, uWS, N)."""
    rx = np.asarray(rx, float); ry = np.asarray(ry, float); ryaw = np.asarray(ryaw, float)
    dt = Ts / sampleN
    rv = np.zeros(rx.size)
    rv[:-1] = np.diff(rx) / dt * np.cos(ryaw[:-1]) + np.diff(ry) / dt * np.sin(ryaw[:-1])          # :222-229
    v, a = velo_smooth(rv, amax, dt)                                                                # :231
    delta = np.arctan(np.diff(ryaw) * L / motionStep * np.sign(v[:-1]))                             # :233
    s = slice(None, None, sampleN)                                                                  # :237-244
    rxs, rys, ryaws, vs = rx[s], ry[s], ryaw[s], v[s]
    a_s, ds = a[s], delta[s]
    N = rxs.size - 1
    xWS = np.column_stack([rxs, rys, ryaws, vs])                                                    # :247
    uWS = np.column_stack([ds, a_s])                                                                # :248  (>= N rows)
    return dict(rx=rxs, ry=rys, ryaw=ryaws, xWS=xWS, uWS=uWS, N=N)


def plan_warm_start(x0, xF, scenario="backwards", Ts=None, L=2.7, sampleN=3):
    """main.jl:215-248 in one call: Hybrid A* from x0 to xF, then the warm-start extraction.  Ts defaults to the scenario's
    variable-time sampling time (main.jl:43-58).  Returns None if the planner finds no path."""
    if Ts is None:
        Ts = (0.6 if scenario == "backwards" else 0.9) / 3 * sampleN
    ox, oy = obstacle_points(scenario)
    rx, ry, ryaw = hybrid_a_star.calc_hybrid_astar_path(float(x0[0]), float(x0[1]), float(x0[2]), float(xF[0]), float(xF[1]), float(xF[2]),
                                                        ox, oy, hybrid_a_star.XY_GRID_RESOLUTION, hybrid_a_star.YAW_GRID_RESOLUTION,
                                                        hybrid_a_star.OB_MAP_RESOLUTION)
    if rx is None:
        return None
    out = warm_start_from_path(rx, ry, ryaw, Ts, L, sampleN, hybrid_a_star.MOTION_RESOLUTION)
    out.update(path=(rx, ry, ryaw), Ts=Ts, ox=ox, oy=oy)
    return out


def _plan_one(args):
    x0, xF, scenario = args
    w = plan_warm_start(x0, xF, scenario)
    if w is not None:
        w.pop("ox", None); w.pop("oy", None)          # the point clouds are per scenario, not per problem
    return w


def plan_batch(x0s, xF, scenario="backwards", workers=None):
    """Hybrid A* warm starts for many start poses (the randomised sweeps of main.jl:165-168), one planner run per pose on a pool
    of host processes.  Returns a list of plan_warm_start() dicts (None where no path was found); N differs from pose to pose,
    as in the reference."""
    import multiprocessing as mp
    jobs = [(np.asarray(x, float), np.asarray(xF, float), scenario) for x in x0s]
    workers = workers or min(len(jobs), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    if workers <= 1 or len(jobs) <= 1:
        return [_plan_one(j) for j in jobs]
    with mp.get_context("spawn").Pool(workers) as pool:      # spawn: safe next to the threads of torch / BLAS in the parent
        return pool.map(_plan_one, jobs, chunksize=max(1, len(jobs) // (4 * workers)))


def group_by_horizon(plans):
    """The batched C-ABI takes one horizon N per call: indices of the successfully planned problems grouped by N."""
    groups = {}
    for i, w in enumerate(plans):
        if w is not None:
            groups.setdefault(int(w["N"]), []).append(i)
    return dict(sorted(groups.items()))
