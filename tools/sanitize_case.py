"""One small solve for compute-sanitizer (tools/sanitize.sh).  python tools/sanitize_case.py parking|dist|quad [B] [max_iter]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, quadcopter, scenarios

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
o = obca_b200.default_opts()
if len(sys.argv) > 3:
    o.max_iter = int(sys.argv[3])
if what == "quad":
    sc = scenarios.quadcopter_batch(B, int(os.environ.get("QUAD_N", "40")), 2)      # N = 40: two warps, N = 100: four (the fast paths)
    r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], sc["N"], sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1, o)
    f, _ = quadcopter.check_quadcopter_batch(r["xp"], r["up"], r["ts"], sc["x0"], sc["xF"], sc["Ts"], r["lp"], sc["obs"], sc["R"])
else:
    sd = 0 if what == "dist" else 1
    sc = scenarios.reverse_parking_batch(B, 80, 3)
    r = parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"], sc["b"],
                                    sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"], sd, None, None, o)
    f, _, _ = parking.check_parking_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"],
                                          sc["b"], r["xp"], r["up"], r["lp"], r["np"], r["ts"], 0, sd, r["sl"])
print(what, "mode", os.environ.get("OBCA_MODE", "0"), "B", B, "exitflag", r["exitflag"].tolist(), "iters", r["iters"].tolist(), "checker", np.asarray(f).tolist())
