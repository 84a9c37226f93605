"""Generate golden solutions with the ORACLE solver (oracle/ipm_ref.py on the restated reference NLP).

Run in the build container:  PYTHONPATH=. python tests/golden/make_golden.py <case> [...]
Each case solves one problem of BASELINE config 2 (reverse parking, N=80, 3 obstacles, numpy default_rng(0) start
poses, deterministic geometric warm start) and stores inputs + the oracle's solution in tests/golden/<case>.npz.
The reference itself (Julia + JuMP + Ipopt) cannot run here, so these are oracle outputs, not reference outputs
(parity unpinned).
"""
import sys
import time

import numpy as np

from obca_b200.scenarios import reverse_parking_batch
from oracle.dualmultws_ref import dualmultws
from oracle.parking_solve import solve_parking

CASES = {
    # name: (problem index, variant, fixTime)            BASELINE config 2 (reverse parking, seed 0)
    "sd_var_p0": (0, "sd", 0),
    "sd_var_p1": (1, "sd", 0),
    "d_var_p0": (0, "d", 0),
    "sd_fix_p2": (2, "sd", 1),
}
# BASELINE config 3 (parallel parking, seed 1; 3 obstacles and the reference's own 4-obstacle list, main.jl:154-157)
CASES_PARALLEL = {
    "par3_sd_p0": (3, 0, "sd", 0),     # name: (n_obstacles, problem index, variant, fixTime)
    "par4_sd_p1": (4, 1, "sd", 0),
    "par4_d_p2": (4, 2, "d", 0),
}
# BASELINE config 4 (QuadcopterSignedDist / QuadcopterDist, N = 100, seed 2)
CASES_QUAD = {
    "quad_sd_p0": (0, "sd"),           # name: (problem index, variant)
    "quad_d_p1": (1, "d"),
}


def main(names):
    sc = reverse_parking_batch(8, 80, 0)
    for name in names:
        if name in CASES_PARALLEL:
            parallel_case(name)
            continue
        if name in CASES_QUAD:
            quad_case(name)
            continue
        i, variant, fix = CASES[name]
        Ts = sc["Ts_fix"] if fix else sc["Ts"]
        t0 = time.time()
        lWS, nWS, d = dualmultws(80, sc["nOb"], sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], sc["ego"])
        out, res, nlp = solve_parking(sc["x0"][i], sc["xF"], 80, Ts, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"],
                                      sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], fix,
                                      sc["xWS"][i], sc["uWS"][i], variant, lWS, nWS)
        xp, up, ts, exitflag, dt, lp, npp = out
        sl = nlp.lay.unpack(res.z)[5]
        print(name, "status", res.status, "iters", res.iters, "err", res.err, "f", nlp.f(res.z), "time", time.time() - t0, flush=True)
        np.savez_compressed(f"tests/golden/{name}.npz", index=i, variant=variant, fixTime=fix, seed=0, N=80,
                            lWS=lWS, nWS=nWS, dWS=d, xp=xp, up=up, ts=ts, lp=lp, np=npp,
                            sl=sl if sl is not None else np.zeros(0), f=nlp.f(res.z), status=res.status,
                            iters=res.iters, err=res.err)


def parallel_case(name):
    from obca_b200.scenarios import parallel_parking_batch
    from oracle import ipm_ref
    nob, i, variant, fix = CASES_PARALLEL[name]
    sc = parallel_parking_batch(8, 80, 1, nob)
    Ts = sc["Ts_fix"] if fix else sc["Ts"]
    t0 = time.time()
    lWS, nWS, d = dualmultws(80, sc["nOb"], sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], sc["ego"])
    out, res, nlp = solve_parking(sc["x0"][i], sc["xF"], 80, Ts, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"],
                                  sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], fix,
                                  sc["xWS"][i], sc["uWS"][i], variant, lWS, nWS, ipm_ref.IpmOptions(linsolve="sparse"))
    xp, up, ts, exitflag, dt, lp, npp = out
    sl = nlp.lay.unpack(res.z)[5]
    print(name, "status", res.status, "iters", res.iters, "err", res.err, "f", nlp.f(res.z), "time", time.time() - t0, flush=True)
    np.savez_compressed(f"tests/golden/{name}.npz", scenario=f"parallel{nob}", index=i, variant=variant, fixTime=fix, seed=1, N=80,
                        lWS=lWS, nWS=nWS, dWS=d, xp=xp, up=up, ts=ts, lp=lp, np=npp,
                        sl=sl if sl is not None else np.zeros(0), f=nlp.f(res.z), status=res.status, iters=res.iters, err=res.err)


def quad_case(name):
    from obca_b200.scenarios import quadcopter_batch
    from oracle import ipm_ref
    from oracle.quadcopter_solve import solve_quadcopter
    i, variant = CASES_QUAD[name]
    N = 100
    sc = quadcopter_batch(4, N, 2)
    t0 = time.time()
    # engine="compiled": oracle/cpu_ipm, the C++ restatement of oracle/ipm_ref.py (same algorithm, same sympy callbacks; the python
    # engine needs ~1 h per N = 100 solve)
    out, res, nlp = solve_quadcopter(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i], 1.0, variant,
                                     ipm_ref.IpmOptions(max_iter=3000), engine="compiled")
    xp, up, ts, exitflag, dt, lp, status = out
    sl = nlp.lay.unpack(res.z)[4]
    print(name, "status", res.status, "exitflag", exitflag, "iters", res.iters, "err", res.err, "f", nlp.f(res.z), "time", time.time() - t0, flush=True)
    np.savez_compressed(f"tests/golden/{name}.npz", scenario="quad", index=i, variant=variant, seed=2, N=N, xp=xp, up=up, ts=ts, lp=lp,
                        sl=sl if sl is not None else np.zeros(0), f=nlp.f(res.z), status=res.status, exitflag=exitflag, iters=res.iters,
                        err=res.err)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
