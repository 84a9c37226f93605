"""Schedules of the parking solver compared and timed (development tool; run under gpurun).

  python tools/gpu_phased.py [B] [quick]

1. B = 200: tail kernel only / rounds only / rounds + hand-over must be bit-identical.
2. B problems (default 4096): device time of the solve for several hand-over points, kernel-time breakdown
   (OBCA_PHASE_TIMING=1), evaluation counters.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
quick = len(sys.argv) > 2


def run(sc, mode, thresh=None, timing=False):
    os.environ["OBCA_MODE"] = mode
    if thresh is None:
        os.environ.pop("OBCA_TAIL_THRESH", None)
    else:
        os.environ["OBCA_TAIL_THRESH"] = str(thresh)
    if timing:
        os.environ["OBCA_PHASE_TIMING"] = "1"
    else:
        os.environ.pop("OBCA_PHASE_TIMING", None)
    return parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                       sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])


def sched():
    lib = obca_b200.lib()
    rnd = C.c_int(0); hand = C.c_int(0); kms = (C.c_double * 5)(); pr = (C.c_ulonglong * 8)()
    lib.obca_last_schedule(C.c_int(0), C.byref(rnd), C.byref(hand), kms)
    lib.obca_last_profile(C.c_int(0), pr)
    return rnd.value, hand.value, [round(x, 3) for x in kms], [int(x) for x in pr]


sc = scenarios.reverse_parking_batch(200, 80, seed=5)
ref = run(sc, "1")
print("B=200 tail: conv", int(ref["exitflag"].sum()), "iters mean", ref["iters"].mean(), "max", ref["iters"].max(), "time ms", ref["time"] * 1e3, flush=True)
for name, mode, th in (("rounds", "2", 0), ("handover", "2", 120), ("auto", "0", None)):
    r = run(sc, mode, th)
    same = all(np.array_equal(r[k], ref[k]) for k in ("xp", "up", "ts", "lp", "np", "iters", "exitflag"))
    d = np.flatnonzero(r["iters"] != ref["iters"])
    print(f"B=200 {name}: conv {int(r['exitflag'].sum())} identical {same}; iters differ at {len(d)} {d[:8].tolist()} "
          f"max|dx| {float(np.abs(r['xp'] - ref['xp']).max()):.3e} time {r['time'] * 1e3:.2f} ms sched {sched()[:2]}", flush=True)

sc = scenarios.reverse_parking_batch(B, 80, 0)
cases = [("0", None), ("2", 0), ("1", None)] if quick else [("0", None), ("2", 0), ("2", 64), ("2", 150), ("2", 300), ("2", 600), ("2", 1200), ("2", 2048), ("1", None)]
for mode, th in cases:
    best = 1e9
    for rep in range(3):
        r = run(sc, mode, th)
        best = min(best, r["time"])
    print(f"B={B} mode {mode} thresh {th}: solve {best * 1e3:.1f} ms -> {B / best:.0f} traj/s; conv {int(r['exitflag'].sum())}; "
          f"iters mean {r['iters'].mean():.1f} max {r['iters'].max()}; sched(rounds, handed) {sched()[:2]}", flush=True)
r = run(sc, "0", None, timing=True)
rnd, hand, kms, pr = sched()
print(f"auto with per-kernel events: rounds {rnd} handed {hand} ms [eval, sweep, step, tail, dualws] = {kms}; total {r['time'] * 1e3:.1f} ms")
print(f"counters: K1 evals rounds {pr[7]} (re-evals {pr[5]}), merit evals rounds {pr[6]}, tail K1 {pr[4]} merit {pr[3]}")
r = run(sc, "2", 0, timing=True)
rnd, hand, kms, pr = sched()
print(f"rounds only with per-kernel events: rounds {rnd} ms [eval, sweep, step, tail, dualws] = {kms}; total {r['time'] * 1e3:.1f} ms; "
      f"per round us: eval {1e3 * kms[0] / rnd:.1f} sweep {1e3 * kms[1] / rnd:.1f} step {1e3 * kms[2] / rnd:.1f}")
print("iteration histogram (bin edges 0..200 step 10):", np.histogram(r["iters"], bins=np.arange(0, 211, 10))[0].tolist())
