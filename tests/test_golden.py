"""Golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py with the ORACLE solver -- the
reference itself cannot run here, parity unpinned).  CPU part: the goldens are KKT points of the restated reference
NLP and pass the verbatim ParkingConstraints; the kernel sources (emulated CTA) land on the same solutions."""
import glob
import os

import numpy as np
import pytest

import emul
from obca_b200 import scenarios
from oracle import checkers, kkt_check

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
               if not os.path.basename(p).startswith(("_", "quad_")))        # quadcopter goldens: tests/test_quadcopter.py


def golden_scenario(g):
    """The batch a golden was generated from (tests/golden/make_golden.py): BASELINE config 2 (reverse parking, seed 0) or
    config 3 (parallel parking, seed 1, 3 or 4 obstacles)."""
    kind = str(g["scenario"]) if "scenario" in g else "reverse"
    if kind.startswith("parallel"):
        return scenarios.parallel_parking_batch(8, 80, int(g["seed"]), int(kind[-1]))
    return scenarios.reverse_parking_batch(8, 80, int(g["seed"]))


def load(case):
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    return g, golden_scenario(g)


@pytest.mark.parametrize("case", CASES)
def test_golden_is_kkt_point_and_feasible(case):
    g, sc = load(case)
    i, variant, fix = int(g["index"]), str(g["variant"]), int(g["fixTime"])
    assert int(g["status"]) == 1
    B = sc["B"]
    r = dict(xp=np.repeat(g["xp"][None], B, 0), up=np.repeat(g["up"][None], B, 0), ts=np.repeat(g["ts"][None], B, 0),
             lp=np.repeat(g["lp"][None], B, 0), np=np.repeat(g["np"][None], B, 0),
             sl=np.repeat(g["sl"][None], B, 0) if variant == "sd" else None)
    e = kkt_check.reference_kkt_error(sc, i, r, variant, fix)
    assert e["E0"] < 1e-5 and abs(e["f"] - float(g["f"])) < 1e-9
    Ts = sc["Ts_fix"] if fix else sc["Ts"]
    ok = checkers.ParkingConstraints(sc["x0"][i], sc["xF"], 80, Ts, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"],
                                     sc["b"], g["xp"], g["up"], g["lp"], g["np"], g["ts"], fix, 1 if variant == "sd" else 0)
    assert ok == 1


@pytest.mark.parametrize("case", CASES)
def test_kernel_sources_reproduce_golden(case):
    g, sc = load(case)
    i, variant, fix = int(g["index"]), str(g["variant"]), int(g["fixTime"])
    lWS = np.repeat(g["lWS"][None], sc["B"], 0); nWS = np.repeat(g["nWS"][None], sc["B"], 0)
    sub = dict(sc); sub.update(B=1, x0=sc["x0"][i:i + 1], rx=sc["rx"][i:i + 1], ry=sc["ry"][i:i + 1], ryaw=sc["ryaw"][i:i + 1],
                               xWS=sc["xWS"][i:i + 1], uWS=sc["uWS"][i:i + 1])
    r = emul.solve_batch(sub, fix, variant, None, lWS[:1], nWS[:1])
    assert r["status"][0] == 1
    # two interior-point runs stopped at tol=1e-5 agree to O(mu) ~ 1e-5 on the primal trajectory
    assert np.abs(r["xp"][0].T - g["xp"]).max() < 2e-4 and np.abs(r["up"][0].T - g["up"]).max() < 2e-4
    assert np.abs(r["ts"][0] - g["ts"]).max() < 1e-5
