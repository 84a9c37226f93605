"""ORACLE (test infrastructure only).  Reference quadcopter NLP (quadcopter_nlp.py) + IPOPT stand-in (ipm_ref.py)."""
from __future__ import annotations

import time

import numpy as np

from . import ipm_ref
from .quadcopter_nlp import build_quadcopter_nlp, initial_point, stage_order


def solve_quadcopter(x0, xF, N, Ts, R, obs, xWS, timeWS=1.0, variant="sd", opts=None, verbose=False, dual_ws=True, engine="python"):
    """13 arguments of QuadcopterSignedDist.jl:25 (ob1..ob5 = obs rows; uWS is ignored by the reference, :202).
    engine: "python" = oracle/ipm_ref.py, "compiled" = its C++ restatement oracle/cpu_ipm (same algorithm, same callbacks)."""
    nlp = build_quadcopter_nlp(x0, xF, N, Ts, R, obs, variant)
    lay = nlp.lay
    for k in (0, N):                    # pinned end states: drop the redundant bounds (same KKT points)
        for j in range(12):
            nlp.zL[lay.x(j, k)] = -np.inf; nlp.zU[lay.x(j, k)] = np.inf
    o = opts or ipm_ref.IpmOptions(max_iter=3000)          # the reference sets no max_iter (Ipopt default 3000, :28-31)
    m = np.zeros(nlp.mE, bool)
    for fam in nlp.eq:
        if fam.name.startswith("norm"):
            m[fam.row0:fam.row0 + fam.n] = True
        if fam.name.startswith("dyn"):
            m[fam.row0 + N - 1] = True
    o.dc_rows = m
    o.freeze_degenerate = 1e-6
    o.verbose = verbose
    if o.linsolve == "sparse":
        o.order = stage_order(nlp)
    xw = np.array(xWS, float)
    xw[:, 0] = x0; xw[:, N] = xF                         # the pinned end states define the first / last warm-start block
    z0 = initial_point(lay, xw, timeWS, obs if dual_ws else None)
    t0 = time.time()
    if engine == "compiled":
        from types import SimpleNamespace
        from . import cpu_ipm
        r = cpu_ipm.solve_batch([nlp], [z0], m, stage_order(nlp), cpu_ipm.default_opts(o), 1)
        res = SimpleNamespace(z=r["z"][0], status=int(r["status"][0]), iters=int(r["iters"][0]), err=float(r["err"][0]))
    else:
        res = ipm_ref.solve(nlp, z0, o)
    xp, up, ts, lp, sl = lay.unpack(res.z)
    exitflag = 1 if res.status == 1 else 0
    if variant == "sd" and exitflag == 1 and sl.sum() > 1e-3:
        exitflag = 2                                       # QuadcopterSignedDist.jl:285-288
    return (xp, up, ts, exitflag, time.time() - t0, lp, "Optimal" if res.status == 1 else "Error"), res, nlp
