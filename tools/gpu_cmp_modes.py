"""Compare OBCA_MODE variants problem by problem (development tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obca_b200 import parking, scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenarios.reverse_parking_batch(B, 80, 0)
def run(mode, thresh=None):
    os.environ["OBCA_MODE"] = mode
    if thresh is None: os.environ.pop("OBCA_TAIL_THRESH", None)
    else: os.environ["OBCA_TAIL_THRESH"] = thresh
    return parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                       sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
res = {"mono": run("3"), "tail": run("1"), "rounds": run("2", "0"), "auto": run("0"), "tail2": run("1")}
ref = res["mono"]
for k, r in res.items():
    d = np.flatnonzero(r["iters"] != ref["iters"])
    dx = np.abs(r["xp"] - ref["xp"]).max(axis=(1, 2))
    print(k, "conv", int(r["exitflag"].sum()), "iters differ at", len(d), "problems", d[:10].tolist(),
          [(int(ref["iters"][i]), int(r["iters"][i])) for i in d[:6]], "max|dx|", float(dx.max()), "n(dx>0)", int((dx > 0).sum()))
