"""Hybrid A* path planner: host-side restatement of AutonomousParking/hybrid_a_star.jl (the producer of the (rx, ry, ryaw)
path the OBCA NLP is warm-started with, main.jl:215-219).  CPU, Python: the reference's planner is
host code as well, and it runs once per problem before the GPU solve.

Search (hybrid_a_star.jl:104-190): best-first over (x, y, yaw) grid cells with continuous states attached to the nodes;
motion primitives = arcs of length XY_GRID_RESOLUTION for N_STEER steering angles each way, forward and backward (:290-298,
:341-393); every popped node first tries the analytic Reeds-Shepp connection to the goal (:193-213, :262-287); heuristic =
grid distance-to-goal with obstacles (:422-426).  Costs: switch-back 10, steering change 10 per rad, backward 0, steering 0 (:60-63).

Known difference: ties in the priority queue are broken by insertion order here (heapq) and by the binary-heap layout of
DataStructures.jl in the reference -- equally good paths may differ.
"""
from __future__ import annotations

import heapq
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
from scipy.spatial import cKDTree

from . import collision, reeds_shepp
from .grid_policy import calc_dist_policy, jround

VEHICLE_RADIUS = 1.0                       # hybrid_a_star.jl:42
OB_MAP_RESOLUTION = 0.1                    # :46
YAW_GRID_RESOLUTION = math.radians(5.0)    # :47
N_STEER = 5.0                              # :48
XY_GRID_RESOLUTION = 0.3                   # :53
MOTION_RESOLUTION = 0.1                    # :54
SB_COST = 10.0                             # :60
BACK_COST = 0.0                            # :61
STEER_CHANGE_COST = 10.0                   # :62
STEER_COST = 0.0                           # :63
H_COST = 1.0                               # :64
WB = 2.7                                   # :66
MAX_STEER = 0.6                            # :67

pi_2_pi = reeds_shepp.pi_2_pi


@dataclass
class Node:                                 # hybrid_a_star.jl:69-80
    xind: int
    yind: int
    yawind: int
    direction: bool
    x: List[float]
    y: List[float]
    yaw: List[float]
    steer: float
    cost: float
    pind: int


@dataclass
class Config:                               # :82-101 (the obstacle-map fields are only used for the bounds here)
    minx: int
    miny: int
    minyaw: int
    maxx: int
    maxy: int
    maxyaw: int
    xw: int
    yw: int
    yaww: int
    xyreso: float
    yawreso: float


def calc_config(ox, oy, xyreso, yawreso) -> Config:          # :455-478
    minx, miny = jround(min(ox) / xyreso), jround(min(oy) / xyreso)
    maxx, maxy = jround(max(ox) / xyreso), jround(max(oy) / xyreso)
    minyaw = jround(-math.pi / yawreso) - 1
    maxyaw = jround(math.pi / yawreso)
    return Config(minx, miny, minyaw, maxx, maxy, maxyaw, maxx - minx, maxy - miny, maxyaw - minyaw, xyreso, yawreso)


def calc_motion_inputs() -> Tuple[List[float], List[float]]:  # :290-298
    step = MAX_STEER / N_STEER
    up = [step * (i + 1) for i in range(int(round(N_STEER)))]
    u = [0.0] + up + [-v for v in up]
    d = [1.0] * len(u) + [-1.0] * len(u)
    return u + u, d


def calc_index(n: Node, c: Config) -> int:                    # :413-419
    return (n.yawind - c.minyaw) * c.xw * c.yw + (n.yind - c.miny) * c.xw + (n.xind - c.minx)


def calc_next_node(cur: Node, c_id: int, u: float, d: float, c: Config) -> Node:     # :341-393
    arc_l = XY_GRID_RESOLUTION
    nlist = jround(arc_l / MOTION_RESOLUTION) + 1
    xl = [0.0] * nlist; yl = [0.0] * nlist; yawl = [0.0] * nlist
    xl[0] = cur.x[-1] + d * MOTION_RESOLUTION * math.cos(cur.yaw[-1])
    yl[0] = cur.y[-1] + d * MOTION_RESOLUTION * math.sin(cur.yaw[-1])
    yawl[0] = pi_2_pi(cur.yaw[-1] + d * MOTION_RESOLUTION / WB * math.tan(u))
    for i in range(nlist - 1):
        xl[i + 1] = xl[i] + d * MOTION_RESOLUTION * math.cos(yawl[i])
        yl[i + 1] = yl[i] + d * MOTION_RESOLUTION * math.sin(yawl[i])
        yawl[i + 1] = pi_2_pi(yawl[i] + d * MOTION_RESOLUTION / WB * math.tan(u))
    direction = d > 0
    added = abs(arc_l) if direction else abs(arc_l) * BACK_COST
    if direction != cur.direction:
        added += SB_COST
    added += STEER_COST * abs(u)
    added += STEER_CHANGE_COST * abs(cur.steer - u)
    return Node(jround(xl[-1] / c.xyreso), jround(yl[-1] / c.xyreso), jround(yawl[-1] / c.yawreso), direction, xl, yl, yawl, u,
                cur.cost + added, c_id)


def calc_rs_path_cost(p: reeds_shepp.Path) -> float:           # :216-259
    cost = 0.0
    for l in p.lengths:
        cost += l if l >= 0 else abs(l) * BACK_COST
    for a, b in zip(p.lengths[:-1], p.lengths[1:]):
        if a * b < 0.0:
            cost += SB_COST
    for ct in p.ctypes:
        if ct != "S":
            cost += STEER_COST * abs(MAX_STEER)
    ul = [(-MAX_STEER if ct == "R" else MAX_STEER if ct == "L" else 0.0) for ct in p.ctypes]
    for a, b in zip(ul[:-1], ul[1:]):
        cost += STEER_CHANGE_COST * abs(b - a)
    return cost


def verify_index(n: Node, c: Config, kdtree, ox, oy) -> bool:  # :301-326
    if not (0 < n.xind - c.minx < c.xw) or not (0 < n.yind - c.miny < c.yw):
        return False
    return collision.check_collision(n.x, n.y, n.yaw, kdtree, ox, oy)


def analytic_expansion(n: Node, goal: Node, kdtree, ox, oy) -> Optional[reeds_shepp.Path]:   # :262-287
    maxc = math.tan(MAX_STEER) / WB
    p = reeds_shepp.calc_shortest_path(n.x[-1], n.y[-1], n.yaw[-1], goal.x[-1], goal.y[-1], goal.yaw[-1], maxc,
                                       step_size=MOTION_RESOLUTION)
    if p is None or not collision.check_collision(p.x, p.y, p.yaw, kdtree, ox, oy):
        return None
    return p


def calc_hybrid_astar_path(sx, sy, syaw, gx, gy, gyaw, ox, oy, xyreso=XY_GRID_RESOLUTION, yawreso=YAW_GRID_RESOLUTION,
                           obreso=OB_MAP_RESOLUTION, max_expansions=200000):
    """hybrid_a_star.jl:104-190.  Returns (rx, ry, ryaw) as numpy arrays sampled every MOTION_RESOLUTION, or (None, None, None)."""
    del obreso            # the reference's fine obstacle map (:481-503) is computed but never consulted by its search
    syaw, gyaw = pi_2_pi(syaw), pi_2_pi(gyaw)
    ox = [float(v) for v in ox]; oy = [float(v) for v in oy]
    c = calc_config(ox, oy, xyreso, yawreso)
    kdtree = cKDTree(np.column_stack([ox, oy]))
    nstart = Node(jround(sx / xyreso), jround(sy / xyreso), jround(syaw / yawreso), True, [sx], [sy], [syaw], 0.0, 0.0, -1)
    ngoal = Node(jround(gx / xyreso), jround(gy / xyreso), jround(gyaw / yawreso), True, [gx], [gy], [gyaw], 0.0, 0.0, -1)
    h_dp, hminx, hminy = calc_dist_policy(gx, gy, ox, oy, xyreso, VEHICLE_RADIUS)       # :422-426

    def cost_of(n: Node) -> float:                                                       # calc_cost, :537-550
        i, j = n.xind - hminx, n.yind - hminy
        h = h_dp[i - 1, j - 1] if (1 <= i <= h_dp.shape[0] and 1 <= j <= h_dp.shape[1]) else math.inf
        return n.cost + H_COST * h

    open_, closed = {}, {}
    sid = calc_index(nstart, c)
    open_[sid] = nstart
    tick = 0
    pq = [(cost_of(nstart), tick, sid)]
    u, d = calc_motion_inputs()
    final = None
    expansions = 0
    while True:
        if not open_ or not pq or expansions > max_expansions:
            return None, None, None
        _, _, c_id = heapq.heappop(pq)
        if c_id not in open_:
            continue
        cur = open_[c_id]
        expansions += 1
        ap = analytic_expansion(cur, ngoal, kdtree, ox, oy)                               # :193-213
        if ap is not None:
            cur.x = cur.x + ap.x[1:-1]; cur.y = cur.y + ap.y[1:-1]; cur.yaw = cur.yaw + ap.yaw[1:-1]
            cur.cost += calc_rs_path_cost(ap)
            closed[calc_index(ngoal, c)] = cur
            final = cur
            break
        del open_[c_id]
        closed[c_id] = cur
        for ui, di in zip(u, d):
            node = calc_next_node(cur, c_id, ui, di, c)
            if not verify_index(node, c, kdtree, ox, oy):
                continue
            nid = calc_index(node, c)
            if nid in closed or nid in open_:
                continue
            open_[nid] = node
            tick += 1
            heapq.heappush(pq, (cost_of(node), tick, nid))
    # get_final_path (:506-534): goal point, then the node chain back to the start cell
    rx, ry, ryaw = list(ngoal.x), list(ngoal.y), list(ngoal.yaw)
    n = final
    while True:
        rx += n.x[::-1]; ry += n.y[::-1]; ryaw += n.yaw[::-1]
        if (n.xind, n.yind, n.yawind) == (nstart.xind, nstart.yind, nstart.yawind):
            break
        n = closed[n.pind]
    return np.array(rx[::-1]), np.array(ry[::-1]), np.array(ryaw[::-1])
