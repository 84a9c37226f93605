"""The randomised sweep of the reference's AutonomousParking/main.jl:165-168 (many start poses, each planned by Hybrid A* and then
solved) as a batch: plan every pose on the host threads (libobca_planner.so), group the problems by the horizon N the planner gave
them (the batched C-ABI takes one N per call), solve each group on the GPU, run the reference's acceptance test on every result.

usage: python examples/sweep_parking.py [B] [backwards|parallel] [seed]           (needs a CUDA device for the solves)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import obca_b200                                            # noqa: E402
from obca_b200 import parking, scenarios                    # noqa: E402
from obca_b200.planner import native, warmstart             # noqa: E402


def sweep(B=256, scenario="backwards", seed=0, signed_dist=1, workers=0):
    sc = scenarios.reverse_parking_scenario() if scenario == "backwards" else scenarios.parallel_parking_scenario(4)
    rng = np.random.default_rng(seed)
    x0 = np.column_stack([rng.uniform(-9, 9, B), rng.uniform(6.8, 9.3, B), np.zeros(B), np.zeros(B)])      # start poses in the lane, v = 0
    ego = np.array([3.7, 1.0, 1.0, 1.0]); XYbounds = np.array([-15.0, 15.0, 1.0, 10.0]); L = 2.7
    t0 = time.time()
    plans = native.plan_batch(x0, sc["xF"], scenario, workers)
    t_plan = time.time() - t0
    groups = warmstart.group_by_horizon(plans)
    out = dict(B=B, planned=sum(w is not None for w in plans), groups=len(groups), t_plan=t_plan, t_solve=0.0, converged=0, feasible=0,
               exitflag=np.zeros(B, int), N=np.array([w["N"] if w else -1 for w in plans]))
    for N, idx in groups.items():
        Ts = plans[idx[0]]["Ts"]
        st = lambda k: np.stack([plans[i][k][:N] if k == "uWS" else plans[i][k] for i in idx])
        r = parking.parking_solve_batch(x0[idx], sc["xF"], N, Ts, L, ego, XYbounds, sc["nOb"], sc["vOb"], sc["A"], sc["b"], st("rx"), st("ry"),
                                        st("ryaw"), 0, st("xWS"), st("uWS"), signed_dist)
        feas, _, _ = parking.check_parking_batch(x0[idx], sc["xF"], N, Ts, L, ego, XYbounds, sc["nOb"], sc["vOb"], sc["A"], sc["b"], r["xp"], r["up"],
                                                 r["lp"], r["np"], r["ts"], 0, signed_dist, r["sl"])
        out["t_solve"] += float(r["time"])
        out["converged"] += int((r["exitflag"] == 1).sum()); out["feasible"] += int(np.asarray(feas).sum())
        out["exitflag"][idx] = r["exitflag"]
    return out


if __name__ == "__main__":
    a = sys.argv
    try:
        o = sweep(int(a[1]) if len(a) > 1 else 256, a[2] if len(a) > 2 else "backwards", int(a[3]) if len(a) > 3 else 0)
    except obca_b200.ObcaError as e:
        print("solve not run:", e); sys.exit(2)
    print(f"{o['B']} start poses: {o['planned']} planned in {o['t_plan']:.2f} s (host threads), {o['groups']} horizons "
          f"N = {o['N'][o['N'] >= 0].min()}..{o['N'].max()}; solved on the GPU in {o['t_solve'] * 1e3:.0f} ms device time: "
          f"{o['converged']} converged, {o['feasible']} pass ParkingConstraints")
