"""Timing of the phase-split driver against the monolithic persistent kernel (development tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenarios.reverse_parking_batch(B, 80, 0)
def run():
    return parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                       sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
ref = None
for mode, thresh in (("3", None), ("1", None), ("2", None), ("2", "0"), ("2", "150"), ("2", "300"), ("2", "1200"), ("2", "2400")):
    os.environ["OBCA_MODE"] = mode
    if thresh is None: os.environ.pop("OBCA_TAIL_THRESH", None)
    else: os.environ["OBCA_TAIL_THRESH"] = thresh
    best = 1e9
    for rep in range(3):
        r = run(); best = min(best, r["time"])
    if ref is None: ref = r
    same = all(np.array_equal(r[k], ref[k]) for k in ("xp", "up", "lp", "np", "iters"))
    print(f"mode {mode} thresh {thresh}: device {best*1e3:.1f} ms -> {B/best:.0f} traj/s; conv {int(r['exitflag'].sum())}; "
          f"iters mean {r['iters'].mean():.1f} max {r['iters'].max()}; identical to mono: {same}", flush=True)
it = ref["iters"]
print("iteration histogram (bin edges 0..200 step 10):", np.histogram(it, bins=np.arange(0, 211, 10))[0].tolist())
