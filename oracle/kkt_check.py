"""ORACLE (test infrastructure only -- never imported by the product path).

Independent KKT certificate in the REFERENCE's formulation.  Given only the primal output of a solver
(xp, up, timeScale, lp, np[, sl] -- what ParkingSignedDist.jl:313 returns) it

  1. evaluates every constraint of the restated reference NLP (oracle/parking_nlp.py) -> constraint violation,
  2. finds multipliers (y free for equalities, z >= 0 for every inequality row / bound) that minimise
     || grad f + J'y - z_L + z_U ||^2 + || gap .* z ||^2   (bounded sparse linear least squares),
  3. reports Ipopt's scaled NLP error E_0 for that primal-dual pair.

Because the multipliers are recomputed here, the certificate does not trust anything the GPU solver says about its
own duals: a small E_0 proves that the returned primal point is a KKT point of the reference NLP.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
from scipy.optimize import lsq_linear

from .parking_nlp import build_parking_nlp


def kkt_certificate(nlp, z, s_max=100.0, act_tol=np.inf, exact=False):
    """act_tol: only inequality rows / bounds whose gap is below it may carry a multiplier.
    exact=False: the sign-constrained least squares is solved as an unconstrained sparse normal-equation solve
    followed by clipping z to >= 0 (fast; the reported residuals are recomputed with the clipped multipliers, so
    the certificate stays valid -- it can only be pessimistic).  exact=True: scipy lsq_linear (slow)."""
    n = nlp.n
    g = nlp.grad(z)
    JE = nlp.JE(z); JI = nlp.JI(z)
    cE = nlp.cE(z); gI = nlp.g(z)
    gL, gU = nlp.gL, nlp.gU
    zL, zU = nlp.zL, nlp.zU
    cols = [JE.T.tocsc()]
    gaps = [np.zeros(nlp.mE)]
    lo = [np.full(nlp.mE, -np.inf)]
    # inequality rows: lower side multiplier enters with -J', upper side with +J'
    iL = np.where(np.isfinite(gL))[0]; iU = np.where(np.isfinite(gU))[0]
    cols.append((-JI[iL]).T.tocsc()); gaps.append(np.maximum(gI[iL] - gL[iL], 0.0)); lo.append(np.zeros(len(iL)))
    cols.append((JI[iU]).T.tocsc()); gaps.append(np.maximum(gU[iU] - gI[iU], 0.0)); lo.append(np.zeros(len(iU)))
    bL = np.where(np.isfinite(zL))[0]; bU = np.where(np.isfinite(zU))[0]
    cols.append(sps.csc_matrix((-np.ones(len(bL)), (bL, np.arange(len(bL)))), shape=(n, len(bL))))
    gaps.append(np.maximum(z[bL] - zL[bL], 0.0)); lo.append(np.zeros(len(bL)))
    cols.append(sps.csc_matrix((np.ones(len(bU)), (bU, np.arange(len(bU)))), shape=(n, len(bU))))
    gaps.append(np.maximum(zU[bU] - z[bU], 0.0)); lo.append(np.zeros(len(bU)))
    M = sps.hstack(cols).tocsc()
    gap = np.concatenate(gaps)
    lo = np.concatenate(lo)
    keep = np.where((lo < 0) | (gap <= act_tol))[0]
    M = M[:, keep].tocsr(); gap = gap[keep]; lo = lo[keep]
    m = M.shape[1]
    Afull = sps.vstack([M, sps.diags(gap)]).tocsr()
    rhs = np.concatenate([-g, np.zeros(m)])
    if exact:
        sol = lsq_linear(Afull, rhs, bounds=(lo, np.full(m, np.inf)), method="trf", lsmr_tol="auto", tol=1e-14,
                         max_iter=200)
        w = sol.x
    else:
        from scipy.sparse.linalg import splu
        Mc = M.tocsc()
        w = np.zeros(m)
        free = np.ones(m, bool)
        for _ in range(8):                      # tiny active-set loop on the sign constraints
            idx = np.where(free)[0]
            Mi = Mc[:, idx]
            Nrm = (Mi.T @ Mi + sps.diags(gap[idx] ** 2 + 1e-14)).tocsc()
            wi = splu(Nrm).solve(-(Mi.T @ g))
            w[:] = 0.0; w[idx] = wi
            neg = (w < lo) & free
            if not neg.any():
                break
            free &= ~neg
        w = np.maximum(w, lo)
    rz = g + M @ w
    comp = gap * w
    viol_I = np.maximum(np.maximum(gL - gI, gI - gU), 0.0)
    viol_B = np.maximum(np.maximum(zL - z, z - zU), 0.0)
    cinf = max(np.abs(cE).max() if nlp.mE else 0.0, viol_I.max() if nlp.mI else 0.0, viol_B.max())
    y = w[lo < 0]; zz = w[lo >= 0]
    sd = max(s_max, (np.abs(y).sum() + zz.sum()) / max(m, 1)) / s_max
    sc = max(s_max, zz.sum() / max(len(zz), 1)) / s_max
    dinf = np.abs(rz).max(); pinf = np.abs(comp).max()
    return dict(E0=max(dinf / sd, cinf, pinf / sc), dual_inf=dinf, constr_viol=cinf, compl=pinf, f=nlp.f(z),
                mult=w)


def reference_kkt_error(sc, i, r, variant="sd", fixTime=0):
    """sc: scenario batch dict (obca_b200.scenarios); r: result dict of parking_solve_batch; i: problem index."""
    N = sc["N"]
    Ts = sc["Ts"] if not fixTime else sc.get("Ts_fix", sc["Ts"])
    xF = np.broadcast_to(np.asarray(sc["xF"], float).reshape(-1, 4), (sc["B"], 4))[i]
    nlp = build_parking_nlp(sc["x0"][i], xF, N, Ts, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"],
                            sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], fixTime, variant)
    z = nlp.lay.pack(r["xp"][i], r["up"][i], r["ts"][i], r["lp"][i], r["np"][i],
                     r["sl"][i] if variant == "sd" else None)
    return kkt_certificate(nlp, z)
