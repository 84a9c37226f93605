"""Holonomic-with-obstacles heuristic of the reference's Hybrid A*: cost-to-go of an 8-connected grid search from the goal
(`a_star.calc_dist_policy`, AutonomousParking/a_star.jl:47-128; the rest of that file, a plain A* path search, is not used by
the parking pipeline).  Cells within `vr` of an obstacle point are blocked (a_star.jl:256-281).

Difference from the reference, on purpose: the reference lowers the cost of a node that is already in its open set without
re-keying its priority-queue entry (a_star.jl:97-102), so its search is not strictly cost-ordered; this restatement is a
textbook Dijkstra (lazy deletion), which returns the exact shortest grid distances -- what the reference's map approximates.
"""
from __future__ import annotations

import heapq
import math

import numpy as np
from scipy.spatial import cKDTree

_MOTION = ((1, 0, 1.0), (0, 1, 1.0), (-1, 0, 1.0), (0, -1, 1.0),
           (-1, -1, math.sqrt(2)), (-1, 1, math.sqrt(2)), (1, -1, math.sqrt(2)), (1, 1, math.sqrt(2)))   # a_star.jl:236-248


def jround(v: float) -> int:
    """Julia's round(Int64, x): ties away from zero."""
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def calc_obstacle_map(ox, oy, reso, vr):
    """a_star.jl:256-281 (ox, oy already divided by reso).  obmap[ix-1, iy-1] <-> grid cell (ix + minx, iy + miny), ix = 1..xw."""
    minx, miny = jround(min(ox)), jround(min(oy))
    maxx, maxy = jround(max(ox)), jround(max(oy))
    xw, yw = maxx - minx, maxy - miny
    tree = cKDTree(np.column_stack([ox, oy]))
    gx, gy = np.meshgrid(np.arange(1, xw + 1) + minx, np.arange(1, yw + 1) + miny, indexing="ij")
    dist, _ = tree.query(np.column_stack([gx.ravel(), gy.ravel()]))
    obmap = (dist <= vr / reso).reshape(xw, yw)
    return obmap, minx, miny, maxx, maxy, xw, yw


def calc_dist_policy(gx, gy, ox, oy, reso, vr):
    """Returns pmap[xw, yw] (np.inf where unreachable); entry [i-1, j-1] belongs to the cell with node index (i + minx, j + miny),
    i.e. the reference's 1-based `pmap[n.x - minx, n.y - miny]` (a_star.jl:118-128), plus (minx, miny)."""
    oxs = [v / reso for v in ox]; oys = [v / reso for v in oy]
    obmap, minx, miny, _, _, xw, yw = calc_obstacle_map(oxs, oys, reso, vr)
    pmap = np.full((xw, yw), np.inf)
    g = (jround(gx / reso), jround(gy / reso))

    def ok(x, y):                                   # verify_node, a_star.jl:209-228
        ix, iy = x - minx, y - miny
        return 0 < ix < xw and 0 < iy < yw and not obmap[ix - 1, iy - 1]

    heap = [(0.0, g[0], g[1])]
    best = {g: 0.0}
    done = set()
    while heap:
        c, x, y = heapq.heappop(heap)
        if (x, y) in done:
            continue
        done.add((x, y))
        ix, iy = x - minx, y - miny
        if 1 <= ix <= xw and 1 <= iy <= yw:
            pmap[ix - 1, iy - 1] = c
        for dx, dy, dc in _MOTION:
            nx, ny = x + dx, y + dy
            if (nx, ny) in done or not ok(nx, ny):
                continue
            nc = c + dc
            if nc < best.get((nx, ny), math.inf):
                best[(nx, ny)] = nc
                heapq.heappush(heap, (nc, nx, ny))
    return pmap, minx, miny
