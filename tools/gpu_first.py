"""First GPU bring-up: correctness against the host emulation + oracle certificate, then a timing sweep."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios
from oracle import kkt_check

out = {}
print("devices", obca_b200.lib().obca_device_count(), flush=True)
ref = np.load(os.path.join(ROOT, "tests/golden/_dev_emul_B64.npz"))
B, N = 64, 80
sc = scenarios.reverse_parking_batch(B, N, 0)
t = time.time()
lp, npp, d = parking.dualmultws_batch(N, 3, sc["vOb"], sc["A"], sc["b"], sc["rx"], sc["ry"], sc["ryaw"], sc["ego"], want_d=True)
print("dualws wall", time.time() - t, "max|lp-emul|", np.abs(lp - ref["lWS"]).max(), "max|np-emul|", np.abs(npp - ref["nWS"]).max(), flush=True)
out["dualws_maxdiff"] = float(max(np.abs(lp - ref["lWS"]).max(), np.abs(npp - ref["nWS"]).max()))
for retry in (0, 1):
    o = obca_b200.default_opts(0, retry)
    t = time.time()
    r = parking.parking_solve_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                    sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"], 1, None, None, o)
    print(f"retry={retry} B=64 wall", time.time() - t, "dev", r["time"], "exit", int(r["exitflag"].sum()), "iters", r["iters"][:16].tolist(), flush=True)
T = lambda a: np.transpose(a, (0, 2, 1))
dx = np.abs(r["xp"] - T(ref["xp"])).max(axis=(1, 2))
print("iters emul", ref["iters"][:16].tolist())
print("max |xp - emul| per problem (first 16)", dx[:16], "overall", dx.max(), flush=True)
out["xp_vs_emul_max"] = float(dx.max()); out["iters_gpu"] = r["iters"].tolist(); out["iters_emul"] = ref["iters"].tolist()
for i in (0, 1, 5):
    e = kkt_check.reference_kkt_error(sc, i, r)
    print("certificate", i, {k: float(v) for k, v in e.items() if k != "mult"}, "solver e0", r["kkt_err"][i], flush=True)
    out[f"cert_{i}"] = float(e["E0"])
feas, e7, strict = parking.check_parking_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                               r["xp"], r["up"], r["lp"], r["np"], r["ts"], 0, 1, r["sl"])
print("ParkingConstraints feasible", int(feas.sum()), "/", B, "strict", int(strict.sum()), "e sums", e7.sum(0).tolist(), flush=True)
# timing sweep
for Bt in (256, 1024, 4096):
    scb = scenarios.reverse_parking_batch(Bt, N, 0)
    for rep in range(2):
        t = time.time()
        rb = parking.parking_solve_batch(scb["x0"], scb["xF"], N, scb["Ts"], scb["L"], scb["ego"], scb["XYbounds"], 3, scb["vOb"], scb["A"], scb["b"],
                                         scb["rx"], scb["ry"], scb["ryaw"], 0, scb["xWS"], scb["uWS"])
        w = time.time() - t
    it = rb["iters"]
    print(f"B={Bt}: wall {w:.3f}s device {rb['time']:.4f}s  -> {Bt/rb['time']:.0f} traj/s (device), converged {int(rb['exitflag'].sum())}/{Bt}, iters mean {it.mean():.1f} max {it.max()}", flush=True)
    out[f"B{Bt}"] = dict(wall=w, dev=rb["time"], conv=int(rb["exitflag"].sum()), it_mean=float(it.mean()), it_max=int(it.max()))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out/first.json"), "w"), indent=1)
