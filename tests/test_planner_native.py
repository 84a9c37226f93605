"""The native warm-start producer (libobca_planner.so, include/obca_planner.h) against the Python restatement of the reference's
planner (obca_b200/planner/*.py): same obstacle clouds, same Reeds-Shepp lengths, same Hybrid A* paths, same warm starts."""
import math
import os
import sys
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from obca_b200.planner import hybrid_a_star, native, reeds_shepp, warmstart      # noqa: E402


def test_symbols_and_obstacle_clouds():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "obca_planner.h")).read()
    names = set(re.findall(r"\b(obca_[a-z_]+)\s*\(", hdr))
    assert {"obca_hybrid_astar", "obca_plan_warmstart", "obca_scenario_obstacle_points", "obca_reeds_shepp_length", "obca_planner_version"} <= names
    for n in names:
        assert hasattr(native.lib(), n), n                              # every entry point the header declares is exported
    assert native.lib().obca_planner_version() == 100
    for sc in ("backwards", "parallel"):
        ox, oy = native.obstacle_points(sc)
        px, py = warmstart.obstacle_points(sc)
        assert np.array_equal(ox, px) and np.array_equal(oy, py)


def test_reeds_shepp_lengths_identical():
    rng = np.random.default_rng(3)
    maxc = math.tan(hybrid_a_star.MAX_STEER) / hybrid_a_star.WB
    for _ in range(300):
        q = [float(v) for v in rng.uniform(-10, 10, 4)]; a = [float(v) for v in rng.uniform(-math.pi, math.pi, 2)]      # Python floats: sum()
        # of numpy scalars is a plain left-to-right sum, of floats a compensated one (CPython >= 3.12) -- the planner works on floats
        lp = reeds_shepp.calc_shortest_path_length(q[0], q[1], a[0], q[2], q[3], a[1], maxc)
        ln = native.reeds_shepp_length(q[0], q[1], a[0], q[2], q[3], a[1], maxc)
        assert lp == ln, (q, a, lp, ln)


@pytest.mark.parametrize("scenario,x0", [("backwards", (-6.0, 9.5, 0.0)), ("backwards", (7.0, 8.0, math.pi)), ("backwards", (3.0, 7.5, 0.3)),
                                         ("parallel", (-6.0, 9.5, 0.0)), ("parallel", (8.0, 8.0, 0.0))])
def test_paths_and_warm_starts_identical(scenario, x0):
    xF = (0.0, 1.3, math.pi / 2) if scenario == "backwards" else (-1.35, 4.0, 0.0)
    ox, oy = warmstart.obstacle_points(scenario)
    t0 = time.time()
    rp = hybrid_a_star.calc_hybrid_astar_path(x0[0], x0[1], x0[2], xF[0], xF[1], xF[2], ox, oy)
    t_py = time.time() - t0
    t0 = time.time()
    rn = native.calc_hybrid_astar_path(x0[0], x0[1], x0[2], xF[0], xF[1], xF[2], ox, oy)
    t_nat = time.time() - t0
    assert rp[0] is not None and rn[0] is not None
    assert rp[0].shape == rn[0].shape
    for a, b in zip(rp, rn):
        assert np.array_equal(a, b)                                     # the same sequence of floating-point operations
    assert t_nat < t_py
    wp = warmstart.plan_warm_start(np.array(x0), np.array(xF), scenario)
    wn = native.plan_warm_start(x0, xF, scenario)
    assert wp["N"] == wn["N"]
    for k in ("rx", "ry", "ryaw"):
        assert np.array_equal(wp[k], wn[k])
    assert np.abs(wp["xWS"] - wn["xWS"]).max() < 1e-12                  # speed: numpy's vectorised cos / sin against libm
    assert np.abs(np.asarray(wp["uWS"])[:wn["N"]] - wn["uWS"]).max() < 1e-12


def test_no_path_and_capacity_codes():
    ox, oy = warmstart.obstacle_points("backwards")
    assert native.calc_hybrid_astar_path(-6.0, 9.5, 0.0, 0.0, 1.3, math.pi / 2, ox, oy, max_expansions=3)[0] is None
    with pytest.raises(native.PlannerError):
        native.calc_hybrid_astar_path(-6.0, 9.5, 0.0, 0.0, 1.3, math.pi / 2, ox, oy, cap=5)


def test_exported_warm_start_files_identical(tmp_path):
    """`export_warmstart` (the hand-over to julia/main_parking.jl) with the native producer and with the Python restatement: the same bytes."""
    from obca_b200.planner import export_warmstart
    a, b = str(tmp_path / "native"), str(tmp_path / "python")
    na = export_warmstart.export(a, "backwards", (-6.0, 9.5, 0.0, 0.0), native=True)
    nb = export_warmstart.export(b, "backwards", (-6.0, 9.5, 0.0, 0.0), native=False)
    assert na == nb > 10
    for f in ("scalars.csv", "path.csv", "xWS.csv", "uWS.csv", "A.csv", "b.csv"):
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f
    c = str(tmp_path / "noplan")
    assert export_warmstart.export(c, "parallel", (-6.0, 9.5, 0.0, 0.0), plan=False) == 0
    assert not os.path.exists(os.path.join(c, "xWS.csv")) and "scenario,1" in open(os.path.join(c, "scalars.csv")).read()


def test_threaded_batch_equals_single_calls():
    rng = np.random.default_rng(5)
    x0s = np.column_stack([rng.uniform(-9, 9, 12), rng.uniform(6.8, 9.3, 12), np.zeros(12)])
    xF = (0.0, 1.3, math.pi / 2)
    t0 = time.time()
    plans = native.plan_batch(x0s, xF, "backwards", workers=4)
    t_batch = time.time() - t0
    assert len(plans) == 12 and all(w is not None for w in plans)
    for i in (0, 5, 11):
        w = native.plan_warm_start(x0s[i], xF, "backwards")
        assert w["N"] == plans[i]["N"]
        for k in ("rx", "ry", "ryaw", "xWS", "uWS"):
            assert np.array_equal(w[k], plans[i][k])
    groups = warmstart.group_by_horizon(plans)                          # the batched C-ABI takes one horizon per call
    assert sum(len(v) for v in groups.values()) == 12 and t_batch < 5.0
