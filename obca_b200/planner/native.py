"""ctypes binding of the native warm-start producer libobca_planner.so (include/obca_planner.h; built in-tree by
__graft_entry__.build() from obca_b200/planner/csrc/obca_planner.cpp).  Same results as the Python modules of this package -- which
stay the documented, line-by-line restatement of the reference -- at about 1/100 of the time."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libobca_planner.so")
_lib = None


class PlannerError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise PlannerError(f"{SO_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(SO_PATH)
        d, i, p = C.c_double, C.c_int, C.c_void_p
        L.obca_planner_version.restype = i
        L.obca_hybrid_astar.restype = i
        L.obca_hybrid_astar.argtypes = [d] * 6 + [p, p, i, d, d, i, i, p, p, p, p]
        L.obca_scenario_obstacle_points.restype = i
        L.obca_scenario_obstacle_points.argtypes = [i, i, p, p, p]
        L.obca_plan_warmstart.restype = i
        L.obca_plan_warmstart.argtypes = [p, p, i, d, d, i, i, p, p, p, p, p, p]
        L.obca_plan_warmstart_batch.restype = i
        L.obca_plan_warmstart_batch.argtypes = [i, p, p, i, d, d, i, i, i] + [p] * 7
        L.obca_reeds_shepp_length.restype = d
        L.obca_reeds_shepp_length.argtypes = [d] * 7
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


SCENARIOS = {"backwards": 0, "parallel": 1}


def obstacle_points(scenario: str):
    n = C.c_int(0)
    ox = np.zeros(1024); oy = np.zeros(1024)
    rc = lib().obca_scenario_obstacle_points(SCENARIOS[scenario], 1024, _p(ox), _p(oy), C.byref(n))
    if rc:
        raise PlannerError(f"obca_scenario_obstacle_points: {rc}")
    return ox[:n.value].copy(), oy[:n.value].copy()


def calc_hybrid_astar_path(sx, sy, syaw, gx, gy, gyaw, ox, oy, xyreso=0.0, yawreso=0.0, max_expansions=200000, cap=8192):
    """(rx, ry, ryaw) sampled every 0.1 m, or (None, None, None) when no path is found (hybrid_a_star.jl:104-190)."""
    ox = np.ascontiguousarray(ox, float); oy = np.ascontiguousarray(oy, float)
    rx = np.zeros(cap); ry = np.zeros(cap); ryaw = np.zeros(cap); n = C.c_int(0)
    rc = lib().obca_hybrid_astar(float(sx), float(sy), float(syaw), float(gx), float(gy), float(gyaw), _p(ox), _p(oy), int(ox.size), float(xyreso),
                                 float(yawreso), int(max_expansions), cap, _p(rx), _p(ry), _p(ryaw), C.byref(n))
    if rc == 1:
        return None, None, None
    if rc:
        raise PlannerError(f"obca_hybrid_astar: {rc}")
    return rx[:n.value].copy(), ry[:n.value].copy(), ryaw[:n.value].copy()


def plan_warm_start(x0, xF, scenario="backwards", Ts=None, L=2.7, sampleN=3, cap=1024):
    """main.jl:215-248 in one native call; same dictionary as obca_b200.planner.warmstart.plan_warm_start (without the full path)."""
    x0 = np.ascontiguousarray(np.asarray(x0, float)[:3]); xF = np.ascontiguousarray(np.asarray(xF, float)[:3])
    rx = np.zeros(cap); ry = np.zeros(cap); ryaw = np.zeros(cap); xWS = np.zeros(4 * cap); uWS = np.zeros(2 * cap); N = C.c_int(0)
    rc = lib().obca_plan_warmstart(_p(x0), _p(xF), SCENARIOS[scenario], float(Ts or 0.0), float(L), int(sampleN), cap, _p(rx), _p(ry), _p(ryaw),
                                   _p(xWS), _p(uWS), C.byref(N))
    if rc == 1:
        return None
    if rc:
        raise PlannerError(f"obca_plan_warmstart: {rc}")
    n = N.value
    Ts = Ts or (0.6 if scenario == "backwards" else 0.9) / 3 * sampleN
    return dict(rx=rx[:n + 1].copy(), ry=ry[:n + 1].copy(), ryaw=ryaw[:n + 1].copy(), xWS=xWS[:4 * (n + 1)].reshape(4, n + 1).T.copy(),
                uWS=uWS[:2 * n].reshape(2, n).T.copy(), N=n, Ts=Ts)


def plan_batch(x0s, xF, scenario="backwards", workers=0, Ts=None, L=2.7, sampleN=3, cap=512):
    """Hybrid A* warm starts for many start poses on a pool of host threads (the randomised sweeps of main.jl:165-168); same list of
    dictionaries as obca_b200.planner.warmstart.plan_batch (None where no path was found)."""
    x0s = np.ascontiguousarray(np.asarray(x0s, float)[:, :3]); B = x0s.shape[0]
    xF = np.ascontiguousarray(np.asarray(xF, float)[:3])
    rx = np.zeros((B, cap)); ry = np.zeros((B, cap)); ryaw = np.zeros((B, cap)); xWS = np.zeros((B, 4 * cap)); uWS = np.zeros((B, 2 * cap))
    N = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
    rc = lib().obca_plan_warmstart_batch(B, _p(x0s), _p(xF), SCENARIOS[scenario], float(Ts or 0.0), float(L), int(sampleN), cap, int(workers),
                                         _p(rx), _p(ry), _p(ryaw), _p(xWS), _p(uWS), _p(N), _p(st))
    if rc or (st > 1).any():
        raise PlannerError(f"obca_plan_warmstart_batch: {rc}, status {st[st > 1][:5]}")
    Ts = Ts or (0.6 if scenario == "backwards" else 0.9) / 3 * sampleN
    out = []
    for i in range(B):
        if st[i]:
            out.append(None); continue
        n = int(N[i])
        out.append(dict(rx=rx[i, :n + 1].copy(), ry=ry[i, :n + 1].copy(), ryaw=ryaw[i, :n + 1].copy(), xWS=xWS[i, :4 * (n + 1)].reshape(4, n + 1).T.copy(),
                        uWS=uWS[i, :2 * n].reshape(2, n).T.copy(), N=n, Ts=Ts))
    return out


def reeds_shepp_length(sx, sy, syaw, gx, gy, gyaw, maxc):
    return lib().obca_reeds_shepp_length(float(sx), float(sy), float(syaw), float(gx), float(gy), float(gyaw), float(maxc))
