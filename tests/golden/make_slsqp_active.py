"""Independent-solver pin with ACTIVE OBCA rows ("missing" #1).

Problem: BASELINE config-2 start pose number 10 of numpy default_rng(0) (x0 = the 11th pose of reverse_parking_batch(16, N, 0)), the
reference's own scenario (main.jl:99-108), horizon shortened to N = 20 (Ts scaled by 80 / N so that the manoeuvre still fits) so
that scipy's SLSQP -- an active-set SQP method that shares nothing with the interior-point codes but the NLP callbacks of
oracle/parking_nlp.py -- finishes in minutes.  At the optimum the car touches its clearance limit: in the Dist variant
d(ego, obstacle) == dmin on 8 (stage, obstacle) blocks, in the SignedDist variant d + sl == dmin on 5 blocks with row
multipliers up to 0.034, i.e. the bilinear (A t - b)' lambda coupling is active, not slack.

SLSQP starts from the reference's initial point (ParkingSignedDist.jl:213-222 with the DualMultWS warm start) after Ipopt's own
projection of the starting point into the bounds (kappa_1 = kappa_2 = 1e-2: lambda, mu >= 0.01 -- at lambda = 0 the rows
|A' lambda|^2 == 1 have a zero gradient and SLSQP's LSQ subproblem is singular).  Output:
tests/golden/_slsqp/slsqp_active_<variant>.npz (z of SLSQP, objective, constraint violation).   Run:  PYTHONPATH=. python
tests/golden/make_slsqp_active.py sd d d_local      (about an hour per variant on 4 cores)
Config 3 (parallel parking, the reference's four obstacles, main.jl:154-162; start pose 1 of parallel_parking_batch(16, N, 1, 4)): tags
p4_sd, p4_d, p4_sd_local, p4_d_local -> slsqp_active_p4_*.npz (one to two minutes each).
  p4_sd_local / p4_d_local   return to the interior-point solutions: primal 3e-5 / 4e-5, lambda and mu on the two active blocks 6e-7
  p4_sd / p4_d               from the far start: other local minima (f = 0.87854 against 0.88066; 21.886 against 21.916), 0.8 m / 2.4 m away

Findings (committed fixtures):
  sd       SLSQP stops (status 8: no further descent at its numerical limit, |c| 3e-14) at f = 6.134674280 -- the interior-point
           solvers stop at 6.134674284: same primal point to 7e-6, lambda on the five active blocks identical, mu to 3e-7.
  d        from the same far start SLSQP heads for a DIFFERENT local minimum (another manoeuvre, f = 21.914 against 22.042 of the
           interior-point solvers, 4.9 m apart; iteration limit reached): the Dist NLP is not convex and which minimum a solver
           reaches depends on its path -- a limit of what any stand-in can say about IPOPT's answer.
  d_local  SLSQP started from the interior-point solution perturbed by 1e-3 (random, seed 0) returns to it: the point
           with 8 active distance rows is a strict local minimiser by an independent method.
"""
import sys
import time

import numpy as np
from scipy.optimize import minimize

from obca_b200 import scenarios
from oracle.dualmultws_ref import dualmultws
from oracle.parking_nlp import build_parking_nlp, initial_point
from oracle.parking_solve import solver_view

N, PROBLEM = 20, 10
PARALLEL_PROBLEM = 1      # config 3 (parallel parking, the reference's four obstacles): start pose 1 of parallel_parking_batch(16, N, 1, 4)


def problem(scenario="reverse"):
    if scenario == "parallel4":
        sc = scenarios.parallel_parking_batch(16, N, 1, 4)
        sc["Ts"] = sc["Ts"] * 80 / N
        return sc, PARALLEL_PROBLEM
    sc = scenarios.reverse_parking_batch(16, N, 0)
    sc["Ts"] = sc["Ts"] * 80 / N
    return sc, PROBLEM


def dense(M):
    return M.toarray() if hasattr(M, "toarray") else np.asarray(M)


def main(variants):
    for tag in variants:
        scenario = "parallel4" if tag.startswith("p4") else "reverse"
        sc, i = problem(scenario)
        parts = tag.split("_")
        variant, local = parts[1] if scenario == "parallel4" else parts[0], tag.endswith("_local")
        nlp = solver_view(build_parking_nlp(sc["x0"][i], sc["xF"], N, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"], sc["b"],
                                            sc["rx"][i], sc["ry"][i], sc["ryaw"][i], 0, variant))
        lay = nlp.lay
        gL, gU = nlp.gL, nlp.gU
        mL, mU = np.isfinite(gL), np.isfinite(gU)
        cons = [dict(type="eq", fun=nlp.cE, jac=lambda z: dense(nlp.JE(z))),
                dict(type="ineq", fun=lambda z: np.concatenate([(nlp.g(z) - gL)[mL], (gU - nlp.g(z))[mU]]),
                     jac=lambda z: np.vstack([dense(nlp.JI(z))[mL], -dense(nlp.JI(z))[mU]]))]
        lWS, nWS, _ = dualmultws(N, sc["nOb"], sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], sc["ego"])
        from oracle.ipm_ref import _push
        z0 = _push(initial_point(lay, sc["xWS"][i], sc["uWS"][i], lWS, nWS), nlp.zL, nlp.zU, 1e-2, 1e-2)
        if local:
            from oracle import cpu_ipm, ipm_ref
            rr = cpu_ipm.ParkingCall(sc, [i], variant, 0).run(opts=cpu_ipm.default_opts(ipm_ref.IpmOptions(tol=1e-9, max_iter=400)),
                                                              lWS=[lWS], nWS=[nWS])
            assert rr["status"][0] == 1
            z0 = rr["z"][0] + 1e-3 * np.random.default_rng(0).normal(size=nlp.n)
            z0 = np.minimum(np.maximum(z0, np.where(np.isfinite(nlp.zL), nlp.zL, -1e300)), np.where(np.isfinite(nlp.zU), nlp.zU, 1e300))
        bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi) for lo, hi in zip(nlp.zL, nlp.zU)]
        t0 = time.time()
        it = [0]

        def cb(z):
            it[0] += 1
            if it[0] % 10 == 0:
                print(f"  {variant} it {it[0]} f {nlp.f(z):.9f} |cE| {np.abs(nlp.cE(z)).max():.2e} t {time.time() - t0:.0f}s", flush=True)
        r = minimize(nlp.f, z0, jac=nlp.grad, method="SLSQP", constraints=cons, bounds=bounds, options=dict(ftol=1e-14, maxiter=600), callback=cb)
        viol = max(np.abs(nlp.cE(r.x)).max(), np.maximum(gL - nlp.g(r.x), 0).max(), np.maximum(nlp.g(r.x) - gU, 0).max())
        print(tag, "status", r.status, r.message, "nit", r.nit, "f", r.fun, "viol", viol, "time", time.time() - t0, flush=True)
        np.savez_compressed(f"tests/golden/_slsqp/slsqp_active_{tag}.npz", z=r.x, f=r.fun, status=r.status, nit=r.nit, viol=viol,
                            N=N, problem=i, variant=variant, lWS=lWS, nWS=nWS)


if __name__ == "__main__":
    main(sys.argv[1:] or ["sd", "d"])
