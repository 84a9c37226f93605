"""One batched solve (development tool for ncu captures).  usage: gpu_one.py B [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obca_b200 import parking, scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = scenarios.reverse_parking_batch(B, 80, 0)
for _ in range(reps):
    r = parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                    sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
print("device ms", r["time"] * 1e3, "conv", int(r["exitflag"].sum()), "iters mean", r["iters"].mean())
