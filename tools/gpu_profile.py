"""Per-phase cycle breakdown of k_parking_solve (clock64 counters) + check against the emulation fixture."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios
ref = np.load(os.path.join(ROOT, "tests/golden/_dev_emul_B64.npz"))
sc = scenarios.reverse_parking_batch(64, 80, 0)
r = parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
T = lambda a: np.transpose(a, (0, 2, 1))
print("B=64 exit", int(r["exitflag"].sum()), "iters", r["iters"][:12].tolist(), "emul", ref["iters"][:12].tolist())
print("max|xp-emul|", np.abs(r["xp"] - T(ref["xp"])).max(), "dev time", r["time"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scb = scenarios.reverse_parking_batch(B, 80, 0)
for rep in range(3):
    rb = parking.parking_solve_batch(scb["x0"], scb["xF"], 80, scb["Ts"], scb["L"], scb["ego"], scb["XYbounds"], 3, scb["vOb"], scb["A"], scb["b"],
                                     scb["rx"], scb["ry"], scb["ryaw"], 0, scb["xWS"], scb["uWS"])
prof = (C.c_ulonglong * 8)()
obca_b200.lib().obca_last_profile(0, prof)
p = np.array(list(prof), float)
names = ["eval_K1", "kkt_K3", "recover", "merit", "update", "serial"]
tot = p[:6].sum()
it = rb["iters"]
print(f"B={B}: device {rb['time']*1e3:.1f} ms -> {B/rb['time']:.0f} traj/s; converged {int(rb['exitflag'].sum())}; iters mean {it.mean():.1f} max {it.max()}")
print("phase share:", {n: round(p[i] / tot, 3) for i, n in enumerate(names)})
print("cycles per problem-iteration:", {n: int(p[i] / it.sum()) for i, n in enumerate(names)}, "total", int(tot / it.sum()))
print("k1 evals / iter", p[7] / it.sum(), "merit evals / iter", p[6] / it.sum())
