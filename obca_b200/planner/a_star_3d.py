"""3-D grid A* of the quadcopter example: host-side restatement of QuadcopterNavigation/a_star_3D.jl:49-265 (the producer of the
position warm start of mainQuadcopter.jl:122-135) plus the point-cloud environment of mainQuadcopter.jl:46-112 (scaled by 10,
as there) and the warm-start assembly of :130-137.

Weighted A* (heuristic weight 1.1, Euclidean) over integer cells with the 26-neighbourhood; a cell is blocked when its nearest
obstacle point is within VEHICLE_RADIUS / reso (a_star_3D.jl:193-231).  Ties in the queue are broken by insertion order (heapq)
here and by the heap layout of Julia's Collections.PriorityQueue in the reference."""
from __future__ import annotations

import heapq
import itertools
import math

import numpy as np
from scipy.spatial import cKDTree

from .grid_policy import jround

VEHICLE_RADIUS = 2.5        # a_star_3D.jl:30
H_WEIGHT = 1.1              # :31
_MOTION = [(dx, dy, dz, math.sqrt(dx * dx + dy * dy + dz * dz))
           for dx, dy, dz in itertools.product((-1, 0, 1), repeat=3) if (dx, dy, dz) != (0, 0, 0)]        # :157-186


def calc_obstacle_map(ox, oy, oz, lo, hi, reso):
    """a_star_3D.jl:193-231.  obmap[ix, iy, iz] <-> cell (ix + minx, iy + miny, iz + minz), ix = 0..xw-1."""
    ox = np.append(np.asarray(ox, float), [lo[0], hi[0]]); oy = np.append(np.asarray(oy, float), [lo[1], hi[1]])
    oz = np.append(np.asarray(oz, float), [lo[2], hi[2]])
    mn = [jround(ox.min()), jround(oy.min()), jround(oz.min())]
    mx = [jround(ox.max()), jround(oy.max()), jround(oz.max())]
    w = [mx[i] - mn[i] for i in range(3)]
    tree = cKDTree(np.column_stack([ox, oy, oz]))
    g = np.stack(np.meshgrid(np.arange(w[0]) + mn[0], np.arange(w[1]) + mn[1], np.arange(w[2]) + mn[2], indexing="ij"), -1).reshape(-1, 3)
    dist, _ = tree.query(g)
    return (dist <= VEHICLE_RADIUS / reso).reshape(w), mn, w


def calc_astar_path(sx, sy, sz, gx, gy, gz, ox, oy, oz, xmin, ymin, zmin, xmax, ymax, zmax, reso):
    """a_star_3D.jl:49-154.  Returns (rx, ry, rz) in the (scaled) units of the inputs, start first."""
    s = (jround(sx / reso), jround(sy / reso), jround(sz / reso))
    g = (jround(gx / reso), jround(gy / reso), jround(gz / reso))
    obmap, mn, w = calc_obstacle_map(np.asarray(ox, float) / reso, np.asarray(oy, float) / reso, np.asarray(oz, float) / reso,
                                     (xmin, ymin, zmin), (xmax, ymax, zmax), reso)
    h = lambda c: math.sqrt((c[0] - g[0]) ** 2 + (c[1] - g[1]) ** 2 + (c[2] - g[2]) ** 2)
    cost = {s: 0.0}
    parent = {s: None}
    closed = set()
    tick = 0
    pq = [(H_WEIGHT * h(s), tick, s)]
    while pq:
        _, _, cur = heapq.heappop(pq)
        if cur in closed:
            continue
        closed.add(cur)
        if cur == g:
            break
        cc = cost[cur]
        for dx, dy, dz, dc in _MOTION:
            n = (cur[0] + dx, cur[1] + dy, cur[2] + dz)
            ix, iy, iz = n[0] - mn[0], n[1] - mn[1], n[2] - mn[2]
            if not (0 < ix < w[0] and 0 < iy < w[1] and 0 < iz < w[2]) or obmap[ix, iy, iz] or n in closed:      # :96-105
                continue
            nc = cc + dc
            if nc < cost.get(n, math.inf):
                cost[n] = nc; parent[n] = cur
                tick += 1
                heapq.heappush(pq, (nc + H_WEIGHT * h(n), tick, n))
    if g not in closed:
        return None, None, None
    path = []
    n = g
    while n is not None:
        path.append(n); n = parent[n]
    p = np.array(path[::-1], float) * reso
    return p[:, 0], p[:, 1], p[:, 2]


def quadcopter_environment():
    """mainQuadcopter.jl:46-112: the two walls as integer point clouds (10 x scale) and the room box."""
    pts = []
    rng = lambda a, b: range(a, b + 1)
    for xx in rng(20, 25):                                   # first wall, with the gap below z = 6
        pts += [(xx, yy, zz) for yy in rng(0, 105) for zz in rng(6, 55)]
    for xx in rng(70, 75):                                   # second wall around the window y in [40, 50], z in [20, 30]
        pts += [(xx, yy, zz) for yy in rng(0, 40) for zz in rng(0, 55)]
        pts += [(xx, yy, zz) for yy in rng(50, 105) for zz in rng(0, 55)]
        pts += [(xx, yy, zz) for yy in rng(40, 50) for zz in rng(30, 55)]
        pts += [(xx, yy, zz) for yy in rng(40, 50) for zz in rng(0, 20)]
    p = np.array(pts, float)
    return p[:, 0], p[:, 1], p[:, 2], (0.0, 0.0, 0.0), (105.0, 105.0, 55.0)


def plan_quadcopter_warm_start(x0, xF, Ts=0.25):
    """mainQuadcopter.jl:114-137: A* on the 10x grid, then xWS (12 x (N+1): positions from the path, everything else zero),
    uWS = 0.5, timeWS = 1, N = path length - 1, Ts scaled so that N Ts stays 80 x the nominal Ts (rounded to 0.01)."""
    ox, oy, oz, lo, hi = quadcopter_environment()
    rx, ry, rz = calc_astar_path(x0[0] * 10.0, x0[1] * 10.0, x0[2] * 10.0, xF[0] * 10.0, xF[1] * 10.0, xF[2] * 10.0, ox, oy, oz,
                                 lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 1.0)
    if rx is None:
        return None
    N = rx.size - 1
    Ts_as = jround(Ts * 80 / N * 100) / 100
    xWS = np.zeros((12, N + 1))
    xWS[0], xWS[1], xWS[2] = rx / 10.0, ry / 10.0, rz / 10.0
    return dict(N=N, Ts=Ts_as, xWS=xWS, uWS=0.5 * np.ones((4, N)), timeWS=1.0, path=(rx, ry, rz))
