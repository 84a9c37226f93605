"""The compiled CPU baseline (oracle/cpu_ipm: the oracle's IPOPT stand-in restated in C++ with generated derivatives and a generic
skyline LDL') against the python oracle it restates and against the golden vectors.  Test infrastructure checking test
infrastructure: bench.py's cpu_baseline / --impl reference legs time this solver."""
import os

import numpy as np
import pytest

from obca_b200 import scenarios
from oracle import cpu_ipm, ipm_ref, kkt_check
from oracle.parking_solve import solve_parking

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("variant,fix", [("sd", 0), ("d", 0), ("sd", 1)])
def test_compiled_solver_tracks_the_python_oracle(variant, fix):
    """Same algorithm, same NLP callbacks (sympy templates, printed as C), same ordering: identical iteration counts and the
    same point to round-off -- the only difference is the factorisation (skyline LDL' vs SuperLU / LAPACK)."""
    N = 20
    sc = scenarios.reverse_parking_batch(3, N, 0)
    sc["Ts"] = sc["Ts"] * 80 / N; sc["Ts_fix"] = sc["Ts_fix"] * 80 / N
    r = cpu_ipm.solve_parking_batch(sc, [0, 2], variant, fix, nthreads=2)
    assert (r["status"] == 1).all()
    for q, i in enumerate([0, 2]):
        lWS = r["dualws"]["z"][q][:r["nlp0"].lay.V * (N + 1)].reshape(N + 1, -1)
        nWS = r["dualws"]["z"][q][r["nlp0"].lay.V * (N + 1):(r["nlp0"].lay.V + 4 * sc["nOb"]) * (N + 1)].reshape(N + 1, -1)
        Ts = sc["Ts_fix"] if fix else sc["Ts"]
        out, res, nlp = solve_parking(sc["x0"][i], sc["xF"], N, Ts, sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                      sc["rx"][i], sc["ry"][i], sc["ryaw"][i], fix, sc["xWS"][i], sc["uWS"][i], variant, lWS, nWS)
        assert res.status == 1 and res.iters == r["iters"][q]
        assert np.abs(res.z - r["z"][q]).max() < 1e-6


def test_compiled_dualmultws_equals_closed_form_distance():
    from oracle.dualmultws_ref import dualmultws
    sc = scenarios.reverse_parking_batch(2, 80, 0)
    c = cpu_ipm.ParkingCall(sc, [1])
    r = c.run()
    d = r["dualws"]["z"][0][c.oD:].reshape(81, 3)
    _, _, dref = dualmultws(80, 3, sc["vOb"], sc["A"], sc["b"], sc["rx"][1], sc["ry"][1], sc["ryaw"][1], sc["ego"])
    assert r["dualws"]["status"][0] == 1 and r["dualws"]["iters"][0] <= 100          # DualMultWS.jl:37 max_iter
    assert np.abs(d - dref).max() < 5e-5


@pytest.mark.parametrize("case", ["sd_var_p0", "d_var_p0", "sd_fix_p2"])
def test_compiled_solver_reproduces_golden(case):
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    i, variant, fix = int(g["index"]), str(g["variant"]), int(g["fixTime"])
    sc = scenarios.reverse_parking_batch(8, 80, int(g["seed"]))
    r = cpu_ipm.ParkingCall(sc, [i], variant, fix).run(lWS=[g["lWS"]], nWS=[g["nWS"]])
    assert r["status"][0] == 1
    xp, up, ts = r["out"][0][:3]
    assert np.abs(xp - g["xp"]).max() < 1e-6 and np.abs(up - g["up"]).max() < 1e-6 and np.abs(ts - g["ts"]).max() < 1e-7
    out = dict(xp=xp[None], up=up[None], ts=ts[None], lp=r["out"][0][3][None], np=r["out"][0][4][None],
               sl=r["out"][0][5][None] if variant == "sd" else None)
    sub = dict(sc); sub.update(B=1, x0=sc["x0"][i:i + 1], rx=sc["rx"][i:i + 1], ry=sc["ry"][i:i + 1], ryaw=sc["ryaw"][i:i + 1])
    e = kkt_check.reference_kkt_error(sub, 0, out, variant, fix)
    assert e["E0"] < 1e-5


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver's contract for `bench.py --impl reference`: one JSON line on stdout (anything a library prints on fd 1 goes to stderr),
    with the keys of the reference arm.  Tiny sample (--cpu-sample) so that the test takes seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "traj/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"].startswith("port") and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
