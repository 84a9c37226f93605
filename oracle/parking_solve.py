"""ORACLE (test infrastructure only).  Glue: reference parking NLP (parking_nlp.py) + IPOPT stand-in (ipm_ref.py).

solve_parking(...) takes the 17 arguments of ParkingSignedDist.jl:29 / ParkingDist.jl:29 and returns the
reference's 7-tuple plus the full solver result.
"""
from __future__ import annotations

import time

import numpy as np

from . import ipm_ref
from .dualmultws_ref import dualmultws
from .parking_nlp import build_parking_nlp, initial_point


def solver_view(nlp):
    """Bounds on the pinned end states are redundant (x[:,1]==x0, x[:,N+1]==xF, ParkingSignedDist.jl:122-131);
    the solver drops them so that no barrier term sits on a fixed variable (same KKT points)."""
    lay = nlp.lay
    for k in (0, lay.N):
        for i in range(4):
            nlp.zL[lay.x(i, k)] = -np.inf; nlp.zU[lay.x(i, k)] = np.inf
    return nlp


def dc_mask(nlp):
    """always-on dual regularisation rows: norm rows (sd) and the terminal dynamics rows."""
    m = np.zeros(nlp.mE, bool)
    N = nlp.lay.N
    for fam in nlp.eq:
        if fam.name.startswith("norm"):
            m[fam.row0:fam.row0 + fam.n] = True
        if fam.name.startswith("dyn"):
            m[fam.row0 + N - 1] = True
    return m


def solve_parking(x0, xF, N, Ts, L, ego, XYbounds, nOb, vOb, A, b, rx, ry, ryaw, fixTime, xWS, uWS,
                  variant="sd", lWS=None, nWS=None, opts=None, verbose=False):
    nlp = solver_view(build_parking_nlp(x0, xF, N, Ts, L, ego, XYbounds, nOb, vOb, A, b, rx, ry, ryaw,
                                        fixTime, variant))
    if lWS is None:
        lWS, nWS, _ = dualmultws(N, nOb, vOb, A, b, rx, ry, ryaw, ego)       # ParkingSignedDist.jl:219
    z0 = initial_point(nlp.lay, xWS, uWS, lWS, nWS)
    o = opts or ipm_ref.IpmOptions()
    o.dc_rows = dc_mask(nlp)
    o.verbose = verbose
    t0 = time.time()
    res = ipm_ref.solve(nlp, z0, o)
    dt = time.time() - t0
    xp, up, ts, lp, np_, sl = nlp.lay.unpack(res.z)
    exitflag = 1 if res.status == 1 else 0
    return (xp, up, ts, exitflag, dt, lp, np_), res, nlp
