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
    # name: (problem index, variant, fixTime)
    "sd_var_p0": (0, "sd", 0),
    "sd_var_p1": (1, "sd", 0),
    "d_var_p0": (0, "d", 0),
    "sd_fix_p2": (2, "sd", 1),
}


def main(names):
    sc = reverse_parking_batch(8, 80, 0)
    for name in names:
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


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
