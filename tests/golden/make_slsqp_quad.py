"""Independent-solver pin for the quadcopter NLPs (QuadcopterSignedDist.jl / QuadcopterDist.jl) with ACTIVE obstacle rows.

Problem: BASELINE config-4 start / goal pair number 5 of numpy default_rng(2) (scenarios.quadcopter_batch(16, N, 2)), the reference's
five boxes, horizon shortened to N = 12 (Ts * N = 20 as in mainQuadcopter.jl:131) so that scipy's SLSQP -- an active-set SQP method that
shares nothing with the interior-point codes but the NLP callbacks of oracle/quadcopter_nlp.py -- finishes in minutes.  At the optimum
four distance rows (ball against box) are active.

  far    SLSQP from the oracle's starting point (reference warm start, closed-form dual start, projected into the bounds like Ipopt does)
  local  SLSQP from the interior-point solution perturbed by 1e-3 (random, seed 0)

Output: tests/golden/_slsqp/slsqp_quad_<variant>_<start>.npz.      Run:  PYTHONPATH=. python tests/golden/make_slsqp_quad.py
"""
import sys
import time

import numpy as np
from scipy.optimize import minimize

from obca_b200 import scenarios
from oracle import ipm_ref
from oracle.quadcopter_nlp import build_quadcopter_nlp, initial_point
from oracle.quadcopter_solve import solve_quadcopter

N, PROBLEM = 12, 5


def dense(M):
    return M.toarray() if hasattr(M, "toarray") else np.asarray(M)


def main(tags):
    sc = scenarios.quadcopter_batch(16, N, 2)
    i = PROBLEM
    for tag in tags:
        variant, start = tag.split("_")
        out, res, nlp = solve_quadcopter(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i], 1.0, variant,
                                         ipm_ref.IpmOptions(tol=1e-9, max_iter=3000), engine="compiled")
        assert res.status == 1
        lay = nlp.lay
        gL, gU = nlp.gL, nlp.gU
        mL, mU = np.isfinite(gL), np.isfinite(gU)
        cons = [dict(type="eq", fun=nlp.cE, jac=lambda z: dense(nlp.JE(z))),
                dict(type="ineq", fun=lambda z: np.concatenate([(nlp.g(z) - gL)[mL], (gU - nlp.g(z))[mU]]),
                     jac=lambda z: np.vstack([dense(nlp.JI(z))[mL], -dense(nlp.JI(z))[mU]]))]
        zL = np.where(np.isfinite(nlp.zL), nlp.zL, -1e300); zU = np.where(np.isfinite(nlp.zU), nlp.zU, 1e300)
        if start == "local":
            z0 = np.clip(res.z + 1e-3 * np.random.default_rng(0).normal(size=nlp.n), zL, zU)
        else:
            xw = np.array(sc["xWS"][i], float); xw[:, 0] = sc["x0"][i]; xw[:, N] = sc["xF"][i]
            z0 = ipm_ref._push(initial_point(lay, xw, 1.0, sc["obs"]), nlp.zL, nlp.zU, 1e-2, 1e-2)
        bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi) for lo, hi in zip(nlp.zL, nlp.zU)]
        t0 = time.time()
        it = [0]

        def cb(z):
            it[0] += 1
            if it[0] % 20 == 0:
                print(f"  {tag} it {it[0]} f {nlp.f(z):.9f} |cE| {np.abs(nlp.cE(z)).max():.2e} t {time.time() - t0:.0f}s", flush=True)
        r = minimize(nlp.f, z0, jac=nlp.grad, method="SLSQP", constraints=cons, bounds=bounds, options=dict(ftol=1e-14, maxiter=1500), callback=cb)
        viol = max(np.abs(nlp.cE(r.x)).max(), np.maximum(gL - nlp.g(r.x), 0).max(), np.maximum(nlp.g(r.x) - gU, 0).max())
        print(tag, "status", r.status, r.message, "nit", r.nit, "f", r.fun, "f_ipm", nlp.f(res.z), "viol", viol, "|z - z_ipm|", np.abs(r.x - res.z).max(),
              "time", time.time() - t0, flush=True)
        np.savez_compressed(f"tests/golden/_slsqp/slsqp_quad_{tag}.npz", z=r.x, f=r.fun, status=r.status, nit=r.nit, viol=viol, N=N, problem=i,
                            variant=variant, z_ipm=res.z, f_ipm=nlp.f(res.z))


if __name__ == "__main__":
    main(sys.argv[1:] or ["sd_local", "d_local", "sd_far", "d_far"])
