"""development: solve time of the default schedule against the hand-over point (best of 4 runs each)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenarios.reverse_parking_batch(B, 80, 0)
os.environ["OBCA_MODE"] = "2"
for th in [int(a) for a in sys.argv[2:]] or [1024, 1536, 2048, 2560, 3072]:
    os.environ["OBCA_TAIL_THRESH"] = str(th)
    best = 1e9
    for rep in range(4):
        r = parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                        sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
        best = min(best, r["time"])
    print(f"B={B} hand-over at {th}: {best * 1e3:.2f} ms", flush=True)
