"""Write everything the reference's main.jl holds right before its two NLP calls (main.jl:215-252) as plain CSV files, so that a
Julia >= 1.6 session can run julia/main_parking.jl (the modernised main.jl: shims of julia/OBCA.jl -> libobca.so) without the
reference's Julia-0.6 planner files:

    python -m obca_b200.planner.export_warmstart OUT_DIR [backwards|parallel] [x0 y0 yaw0] [--native] [--no-plan]

OUT_DIR/scalars.csv   name,value: N, Ts, L, fixTime, nOb
OUT_DIR/x0.csv xF.csv ego.csv XYbounds.csv vOb.csv (one row each), A.csv (sum(vOb) x 2), b.csv (sum(vOb) x 1)
OUT_DIR/path.csv      rx,ry,ryaw   (N+1 rows: the down-sampled Hybrid A* path, main.jl:237-239)
OUT_DIR/xWS.csv       (N+1) x 4,   OUT_DIR/uWS.csv  N x 2   (main.jl:247-248)
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .. import scenarios
from . import warmstart


def export(out_dir, scenario="backwards", x0=(-6.0, 9.5, 0.0, 0.0), plan=True, native=False):
    """plan=False: scenario files only (julia/main_parking.jl then plans with libobca_planner.so); native=True: plan with the native
    producer instead of the Python restatement (same path, tests/test_planner_native.py)."""
    sc = scenarios.reverse_parking_scenario() if scenario == "backwards" else scenarios.parallel_parking_scenario(4)
    x0 = np.asarray(x0, float)
    os.makedirs(out_dir, exist_ok=True)
    sv = lambda name, a: np.savetxt(os.path.join(out_dir, name), np.atleast_2d(np.asarray(a, float)), delimiter=",", fmt="%.17g")
    Ts0 = 0.6 if scenario == "backwards" else 0.9
    if not plan:
        with open(os.path.join(out_dir, "scalars.csv"), "w") as f:
            f.write(f"N,0\nTs,{Ts0:.17g}\nL,2.7\nfixTime,0\nnOb,{sc['nOb']}\nscenario,{0 if scenario == 'backwards' else 1}\n")
        sv("x0.csv", x0); sv("xF.csv", sc["xF"]); sv("ego.csv", [3.7, 1.0, 1.0, 1.0]); sv("XYbounds.csv", [-15.0, 15.0, 1.0, 10.0])
        sv("vOb.csv", sc["vOb"]); sv("A.csv", sc["A"]); sv("b.csv", np.asarray(sc["b"]).reshape(-1, 1))
        return 0
    if native:
        from . import native as nat
        w = nat.plan_warm_start(x0, sc["xF"], scenario)
    else:
        w = warmstart.plan_warm_start(x0, sc["xF"], scenario)
    if w is None:
        raise RuntimeError("Hybrid A*: no path found")
    N = w["N"]
    with open(os.path.join(out_dir, "scalars.csv"), "w") as f:
        f.write(f"N,{N}\nTs,{w['Ts']:.17g}\nL,2.7\nfixTime,0\nnOb,{sc['nOb']}\nscenario,{0 if scenario == 'backwards' else 1}\n")
    sv("x0.csv", x0); sv("xF.csv", sc["xF"]); sv("ego.csv", [3.7, 1.0, 1.0, 1.0]); sv("XYbounds.csv", [-15.0, 15.0, 1.0, 10.0])
    sv("vOb.csv", sc["vOb"]); sv("A.csv", sc["A"]); sv("b.csv", np.asarray(sc["b"]).reshape(-1, 1))
    np.savetxt(os.path.join(out_dir, "path.csv"), np.stack([w["rx"], w["ry"], w["ryaw"]], 1), delimiter=",", fmt="%.17g", header="rx,ry,ryaw",
               comments="")
    sv("xWS.csv", w["xWS"]); sv("uWS.csv", np.asarray(w["uWS"])[:N])
    return N


if __name__ == "__main__":
    flags = [v for v in sys.argv if v.startswith("--")]
    a = [v for v in sys.argv if not v.startswith("--")]
    n = export(a[1], a[2] if len(a) > 2 else "backwards", [float(v) for v in a[3:6]] + [0.0] if len(a) >= 6 else (-6.0, 9.5, 0.0, 0.0),
               plan="--no-plan" not in flags, native="--native" in flags)
    print(f"warm start with N = {n} written to {a[1]}")
