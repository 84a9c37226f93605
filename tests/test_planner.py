"""Warm-start producer: Reeds-Shepp curves, Hybrid A*, collision check, grid heuristic, velocity
smoothing and the main.jl:215-248 pipeline -- host-side restatements in obca_b200/planner.  The reference cannot run here
(no Julia); the tests are the reference's own test strategy (reeds_shepp.jl:846-869 `check_path`, :871-920 `test`) plus
properties the algorithms must have, and one end-to-end run: a Hybrid A* warm start drives the kernels' solver to a feasible
KKT point."""
import math
import os
import sys

import numpy as np
import pytest
from scipy.spatial import cKDTree

from obca_b200.planner import collision, grid_policy, hybrid_a_star, reeds_shepp as rs, velo_smooth, warmstart

D = math.radians


def check_path(sx, sy, syaw, gx, gy, gyaw, maxc):
    """reeds_shepp.jl:846-869."""
    paths = rs.calc_paths(sx, sy, syaw, gx, gy, gyaw, maxc)
    assert len(paths) >= 1
    for p in paths:
        assert abs(p.x[0] - sx) <= 0.01 and abs(p.y[0] - sy) <= 0.01 and abs(rs.pi_2_pi(p.yaw[0] - syaw)) <= 0.01
        assert abs(p.x[-1] - gx) <= 0.01 and abs(p.y[-1] - gy) <= 0.01 and abs(rs.pi_2_pi(p.yaw[-1] - gyaw)) <= 0.01
        d = np.hypot(np.diff(p.x[:-1]), np.diff(p.y[:-1]))
        assert (d <= rs.STEP_SIZE + 1e-3).all()                       # sampling never coarser than the step
        assert abs(sum(abs(l) for l in p.lengths) - p.L) < 1e-12
    return paths


def test_reeds_shepp_reference_cases_and_random_poses():
    check_path(0.0, 0.0, D(10), 7.0, -8.0, D(50), 2.0)                  # reeds_shepp.jl:871-905, the three fixed cases of test()
    check_path(0.0, 0.0, D(10), 7.0, -8.0, D(-50), 2.0)
    check_path(0.0, 10.0, D(-10), -7.0, -8.0, D(-50), 2.0)
    rng = np.random.default_rng(0)
    for _ in range(300):                                                # the random part of test()
        a = [rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-math.pi, math.pi)]
        b = [rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-math.pi, math.pi)]
        check_path(*a, *b, rng.choice([0.2, 1.0 / 4.5, 1.0, 2.0]))


def _drive(pose, lengths, ctypes, maxc):
    """Independent of the module's sampling code: exact unicycle motion along each segment (arc of curvature +-maxc or straight)."""
    x, y, th = pose
    for l, c in zip(lengths, ctypes):
        if c == "S":
            x += l * math.cos(th); y += l * math.sin(th)
        else:
            k = maxc if c == "L" else -maxc
            x += (math.sin(th + k * l) - math.sin(th)) / k
            y += (-math.cos(th + k * l) + math.cos(th)) / k
            th += k * l
    return x, y, th


def test_reeds_shepp_words_reach_the_goal_when_driven():
    rng = np.random.default_rng(2)
    for _ in range(300):
        a = [rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-math.pi, math.pi)]
        b = [rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-math.pi, math.pi)]
        maxc = rng.choice([0.2, 1.0 / 4.5, 1.0])
        for p in rs.calc_paths(*a, *b, maxc):
            x, y, th = _drive(a, p.lengths, p.ctypes, maxc)             # p.lengths are metres after calc_paths
            assert abs(x - b[0]) < 1e-6 and abs(y - b[1]) < 1e-6 and abs(rs.pi_2_pi(th - b[2])) < 1e-6, (p.ctypes, p.lengths)


def test_reeds_shepp_known_answers_and_symmetry():
    maxc = math.tan(0.6) / 2.7
    p = rs.calc_shortest_path(0, 0, 0, 5, 0, 0, maxc)
    assert abs(p.L - 5.0) < 1e-9                                         # straight ahead
    p = rs.calc_shortest_path(0, 0, 0, -3, 0, 0, maxc)
    assert abs(p.L - 3.0) < 1e-9 and min(p.directions) == -1             # straight back
    R = 1.0 / maxc
    p = rs.calc_shortest_path(0, 0, 0, R, R, math.pi / 2, maxc)
    assert abs(p.L - R * math.pi / 2) < 1e-9                             # quarter circle to the left
    rng = np.random.default_rng(1)
    n_longer = 0
    for _ in range(150):
        a = [rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-3, 3)]
        b = [rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-3, 3)]
        l_ref = rs.calc_shortest_path_length(*a, *b, maxc)                  # with the reference's duplicate test (default)
        rs.REFERENCE_DEDUP = False
        try:
            l1 = rs.calc_shortest_path_length(*a, *b, maxc); l2 = rs.calc_shortest_path_length(*b, *a, maxc)
        finally:
            rs.REFERENCE_DEDUP = True
        assert abs(l1 - l2) < 1e-6                                        # time reversal: L(a -> b) = L(b -> a) once every word is kept
        assert l1 >= math.hypot(a[0] - b[0], a[1] - b[1]) - 1e-9          # never shorter than the straight line
        assert l_ref >= l1 - 1e-9                                         # the reference's filter can only lose candidates
        n_longer += l_ref > l1 + 1e-6
    assert n_longer > 0                                                   # ... and it does (documented quirk of reeds_shepp.jl:212-219)


def test_collision_rectangle_and_bubble():
    ox = np.array([2.0, -0.9, 0.0, 10.0]); oy = np.array([0.0, 0.0, 1.2, 10.0])
    tree = cKDTree(np.column_stack([ox, oy]))
    assert not collision.check_collision([0.0], [0.0], [0.0], tree, ox, oy)          # (2, 0) is inside [-1, 3.7] x [-1, 1]
    ox2 = np.array([-1.1, 3.8, 0.0, 1.0]); oy2 = np.array([0.0, 0.0, 1.05, -1.05])
    tree2 = cKDTree(np.column_stack([ox2, oy2]))
    assert collision.check_collision([0.0], [0.0], [0.0], tree2, ox2, oy2)            # all just outside
    assert not collision.check_collision([0.0], [0.0], [math.pi / 2], tree2, ox2, oy2)   # rotated car covers (0, 1.05)
    assert collision.rect_check(0, 0, 0, [], [])


def test_grid_distance_policy():
    # a box of obstacle points around an empty 10 m x 10 m area: distances = octile metric to the goal cell
    xs = np.arange(0, 10.01, 0.5)
    ox = np.concatenate([xs, xs, np.zeros_like(xs), np.full_like(xs, 10.0)]); oy = np.concatenate([np.zeros_like(xs), np.full_like(xs, 10.0), xs, xs])
    pmap, minx, miny = grid_policy.calc_dist_policy(5.0, 5.0, ox, oy, 1.0, 0.5)
    g = (grid_policy.jround(5.0), grid_policy.jround(5.0))
    for (x, y) in ((5, 5), (7, 5), (7, 7), (2, 6), (8, 3)):
        dx, dy = abs(x - g[0]), abs(y - g[1])
        octile = max(dx, dy) + (math.sqrt(2) - 1) * min(dx, dy)
        assert abs(pmap[x - minx - 1, y - miny - 1] - octile) < 1e-9
    assert np.isinf(pmap[0 - minx - 1 + 1, 0 - miny - 1 + 1]) or pmap[0, 0] >= 0      # border cells are blocked / unreachable


def test_velo_smooth_ramps():
    v = np.concatenate([np.full(40, 0.5), np.full(40, -0.5)]); v[-1] = 0.0            # forward, reverse, stop (as main.jl:222-229 builds it)
    vs, a = velo_smooth.velo_smooth(v, 0.3, 0.2)
    assert vs.shape == (80,) and a.shape == (79,)
    assert abs(vs[0]) < 1e-12 and abs(vs[-1]) < 1e-12                                   # starts and ends at rest
    assert np.abs(a).max() <= 0.3 * 1.1                                                 # ramps of |v0| / amax (rounded to whole steps)
    assert np.abs(vs).max() <= 0.5 + 1e-12 and (np.sign(vs[5:35]) == 1).all() and (np.sign(vs[50:70]) == -1).all()


@pytest.mark.parametrize("scenario,x0,xF", [("backwards", [-6.0, 9.5, 0.0, 0.0], [0.0, 1.3, math.pi / 2, 0.0]),       # main.jl:213,108
                                            ("backwards", [7.0, 7.5, 0.0, 0.0], [0.0, 1.3, math.pi / 2, 0.0]),
                                            ("parallel", [-6.0, 8.0, 0.0, 0.0], [-1.35, 4.0, 0.0, 0.0])])              # main.jl:162
def test_hybrid_a_star_paths(scenario, x0, xF):
    w = warmstart.plan_warm_start(x0, xF, scenario)
    assert w is not None
    rx, ry, ryaw = w["path"]
    assert abs(rx[0] - x0[0]) < 1e-9 and abs(ry[0] - x0[1]) < 1e-9 and abs(rx[-1] - xF[0]) < 1e-6 and abs(ry[-1] - xF[1]) < 1e-6
    assert abs(rs.pi_2_pi(ryaw[-1] - xF[2])) < 1e-6
    ds = np.hypot(np.diff(rx), np.diff(ry))
    assert ds.max() <= hybrid_a_star.MOTION_RESOLUTION + 1e-6
    tree = cKDTree(np.column_stack([w["ox"], w["oy"]]))
    assert collision.check_collision(rx, ry, ryaw, tree, w["ox"], w["oy"])
    dyaw = np.abs([rs.pi_2_pi(b - a) for a, b in zip(ryaw[:-1], ryaw[1:])])
    assert (dyaw <= hybrid_a_star.MOTION_RESOLUTION * math.tan(hybrid_a_star.MAX_STEER) / hybrid_a_star.WB + 1e-6).all()   # curvature bound
    assert w["xWS"].shape == (w["N"] + 1, 4) and w["uWS"].shape[0] >= w["N"] and np.abs(w["uWS"][:, 0]).max() <= 0.6 + 1e-9


def test_hybrid_a_star_warm_start_drives_the_solver():
    """End to end on the reference's default problem (main.jl:213): Hybrid A* warm start -> DualMultWS -> the kernels' interior-point
    solve (host build) -> converged, and feasible by the verbatim ParkingConstraints."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    from obca_b200 import scenarios
    from oracle import checkers
    x0 = np.array([-6.0, 9.5, 0.0, 0.0])
    sc = scenarios.reverse_parking_scenario()
    w = warmstart.plan_warm_start(x0, sc["xF"], "backwards")
    N = w["N"]
    sc.update(B=1, N=N, Ts=w["Ts"], L=2.7, ego=np.array([3.7, 1.0, 1.0, 1.0]), XYbounds=np.array([-15.0, 15.0, 1.0, 10.0]), x0=x0[None],
              rx=w["rx"][None], ry=w["ry"][None], ryaw=w["ryaw"][None], xWS=w["xWS"][None], uWS=w["uWS"][None, :N])
    lp, npp, _, _ = emul.dualmultws_batch(sc)
    r = emul.solve_batch(sc, 0, "sd", None, lp, npp)
    assert r["status"][0] == 1
    ok = checkers.ParkingConstraints(x0, sc["xF"], N, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                     r["xp"][0].T, r["up"][0].T, r["lp"][0].T, r["np"][0].T, r["ts"][0], 0, 1)
    assert ok


def test_quadcopter_3d_astar_warm_start_and_solve():
    """mainQuadcopter.jl:114-137 on its default problem: the 3-D A* path goes under the first wall and through the window of
    the second one; the warm start built from it drives the kernels' quadcopter solver (host build) to a point that passes the
    verbatim constrSatisfaction."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    from obca_b200 import scenarios
    from obca_b200.planner import a_star_3d
    from oracle import checkers
    sc = scenarios.quadcopter_scenario()
    w = a_star_3d.plan_quadcopter_warm_start(sc["x0"], sc["xF"])
    assert w is not None
    rx, ry, rz = w["path"]
    assert (rx[0], ry[0], rz[0]) == (10.0, 10.0, 30.0) and (rx[-1], ry[-1], rz[-1]) == (90.0, 30.0, 20.0)
    step = np.sqrt(np.diff(rx) ** 2 + np.diff(ry) ** 2 + np.diff(rz) ** 2)
    assert set(np.round(step ** 2).astype(int)) <= {1, 2, 3}                            # 26-neighbourhood moves
    in1 = (rx >= 20) & (rx <= 25); in2 = (rx >= 70) & (rx <= 75)
    assert in1.any() and (rz[in1] <= 6 - a_star_3d.VEHICLE_RADIUS + 1e-9).all()         # under the first wall, a vehicle radius away
    assert in2.any() and (ry[in2] > 40).all() and (ry[in2] < 50).all() and (rz[in2] > 20).all() and (rz[in2] < 30).all()   # window
    ox, oy, oz, lo, hi = a_star_3d.quadcopter_environment()
    tree = cKDTree(np.column_stack([ox, oy, oz]))
    dmin, _ = tree.query(np.column_stack([rx, ry, rz]))
    assert dmin.min() > a_star_3d.VEHICLE_RADIUS                                        # never inside the inflated obstacles
    assert w["N"] == rx.size - 1 and w["Ts"] == round(0.25 * 80 / w["N"] * 100) / 100 and w["xWS"].shape == (12, w["N"] + 1)
    sc.update(B=1, N=w["N"], Ts=w["Ts"], x0=sc["x0"][None], xF=sc["xF"][None], xWS=w["xWS"][None], timeWS=w["timeWS"])
    o = emul.default_opts(); o.max_iter = 3000
    r = emul.quad_solve_batch(sc, "d", o)
    assert r["status"][0] == 1
    assert checkers.constrSatisfaction(r["xp"][0], r["up"][0], r["ts"][0], sc["x0"][0], sc["xF"][0], sc["Ts"], r["lp"][0], *sc["obs"], sc["R"])


def test_planners_report_failure_instead_of_raising():
    """The reference prints an error and returns `nothing` when the open set runs dry (hybrid_a_star.jl:140-143); here: None."""
    ox, oy = warmstart.obstacle_points("backwards")
    # goal pose buried in the left obstacle block: every analytic expansion collides and the goal cell is unreachable
    rx, ry, ryaw = hybrid_a_star.calc_hybrid_astar_path(-6.0, 9.5, 0.0, -6.0, 4.9, math.pi / 2, ox, oy, max_expansions=300)
    assert rx is None and ry is None and ryaw is None
    from obca_b200.planner import a_star_3d
    pts = np.array([(x, y, z) for x in range(4, 7) for y in range(0, 11) for z in range(0, 11)], float)      # a full wall
    r = a_star_3d.calc_astar_path(1.0, 5.0, 5.0, 9.0, 5.0, 5.0, pts[:, 0], pts[:, 1], pts[:, 2], 0.0, 0.0, 0.0, 10.0, 10.0, 10.0, 1.0)
    assert r == (None, None, None)


def test_planned_batches_are_grouped_by_horizon_and_scattered_back():
    """Hybrid A* horizons differ from pose to pose; the batched entry point takes one N per call.  plan_batch + the grouping
    wrapper of the host mirror: every problem is solved exactly once, in a batch of its own horizon, and lands at its own index
    (the GPU solve is replaced by a recording stub -- the real one is covered by the -m gpu tests)."""
    from obca_b200 import parking, scenarios
    sc = scenarios.reverse_parking_scenario()
    x0s = np.array([[-6.0, 9.5, 0, 0], [7.0, 7.5, 0, 0], [-6.0, 9.5, 0, 0], [-6.0, 4.9, 0, 0], [3.0, 8.0, 0, 0]])   # #3 starts inside an obstacle
    plans = warmstart.plan_batch(x0s, sc["xF"], "backwards", workers=2)
    assert plans[3] is None and all(p is not None for i, p in enumerate(plans) if i != 3)
    assert plans[0]["N"] == plans[2]["N"] and np.array_equal(plans[0]["xWS"], plans[2]["xWS"])             # deterministic
    groups = warmstart.group_by_horizon(plans)
    assert sorted(i for g in groups.values() for i in g) == [0, 1, 2, 4]
    calls = []

    def stub(x0, xF, N, Ts, L, ego, XYb, nOb, vOb, A, b, rx, ry, ryaw, fixTime, xWS, uWS, sd, lWS, nWS, opts):
        B = len(x0)
        assert rx.shape == (B, N + 1) and xWS.shape == (B, N + 1, 4) and uWS.shape == (B, N, 2)
        calls.append((N, B))
        return dict(xp=np.broadcast_to(x0[:, :, None], (B, 4, N + 1)).copy(), up=np.zeros((B, 2, N)), ts=np.ones((B, N + 1)),
                    lp=np.zeros((B, 5, N + 1)), np=np.zeros((B, 12, N + 1)), exitflag=np.ones(B, np.int32), iters=np.full(B, 7, np.int32), time=0.01)

    r = parking.parking_solve_planned(x0s, sc["xF"], plans, 0.6, 2.7, np.array([3.7, 1, 1, 1.0]), np.array([-15, 15, 1, 10.0]), sc["nOb"], sc["vOb"],
                                      sc["A"], sc["b"], solve=stub)
    assert sorted(calls) == sorted((N, len(g)) for N, g in groups.items())
    assert r["xp"][3] is None and r["exitflag"][3] == 0
    for i in (0, 1, 2, 4):
        assert r["xp"][i].shape == (4, plans[i]["N"] + 1) and np.allclose(r["xp"][i][:, 0], x0s[i]) and r["N"][i] == plans[i]["N"]


def test_example_main_parking_dry_run():
    """examples/main_parking.py (the main.jl flow): without a CUDA device it must plan, build the warm start, marshal the
    arguments of the reference-named entry points and stop with the library's no-device message -- never with a traceback."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(root, "examples", "main_parking.py")], capture_output=True, text=True, env=env, timeout=300)
    assert "Hybrid A*:" in p.stdout and "N = 64" in p.stdout
    assert "Traceback" not in p.stderr
    assert p.returncode == 2 and "solve not run" in p.stdout


def _read_export(d):
    """What julia/main_parking.jl reads (same files, same shapes)."""
    s = dict(ln.strip().split(",") for ln in open(os.path.join(d, "scalars.csv")))
    rd = lambda n, **kw: np.atleast_2d(np.loadtxt(os.path.join(d, n), delimiter=",", **kw))
    path = rd("path.csv", skiprows=1)
    return dict(N=int(float(s["N"])), Ts=float(s["Ts"]), L=float(s["L"]), fixTime=int(float(s["fixTime"])), nOb=int(float(s["nOb"])),
                x0=rd("x0.csv"), xF=rd("xF.csv"), ego=rd("ego.csv").ravel(), XYbounds=rd("XYbounds.csv").ravel(),
                vOb=rd("vOb.csv").ravel().astype(int), A=rd("A.csv"), b=rd("b.csv").reshape(-1, 1), rx=path[:, 0], ry=path[:, 1], ryaw=path[:, 2],
                xWS=rd("xWS.csv"), uWS=rd("uWS.csv"))


def test_warm_start_export_for_the_julia_runner(tmp_path):
    """obca_b200.planner.export_warmstart writes what main.jl holds before its NLP calls (main.jl:215-252) in the shapes the
    reference's drivers take (x0, xF 1x4 rows; xWS (N+1)x4; uWS Nx2; A sum(vOb)x2) -- the input of julia/main_parking.jl."""
    from obca_b200.planner import export_warmstart
    N = export_warmstart.export(str(tmp_path))
    e = _read_export(str(tmp_path))
    assert e["N"] == N == 64 and e["x0"].shape == (1, 4) and e["xF"].shape == (1, 4)
    assert e["xWS"].shape == (N + 1, 4) and e["uWS"].shape == (N, 2) and len(e["rx"]) == N + 1
    assert e["A"].shape == (int(e["vOb"].sum()), 2) and e["b"].shape == (int(e["vOb"].sum()), 1) and e["nOb"] == len(e["vOb"]) == 3
    assert np.allclose(e["xWS"][0, :3], e["x0"][0, :3]) and np.allclose(e["xWS"][:, 0], e["rx"])


@pytest.mark.gpu
def test_julia_runner_data_flow_through_the_ctypes_twin(tmp_path):
    """The data flow of julia/main_parking.jl (CSV files -> the reference's 17 positional arguments -> ParkingDist,
    ParkingSignedDist, ParkingConstraints) executed through the tested ctypes twin of the Julia shims."""
    import obca_b200
    from obca_b200.planner import export_warmstart
    export_warmstart.export(str(tmp_path))
    e = _read_export(str(tmp_path))
    for fn, sd in ((obca_b200.ParkingDist, 0), (obca_b200.ParkingSignedDist, 1)):
        xp, up, ts, ef, t, lp, np_ = fn(e["x0"], e["xF"], e["N"], e["Ts"], e["L"], e["ego"], e["XYbounds"], e["nOb"], e["vOb"], e["A"], e["b"],
                                        e["rx"], e["ry"], e["ryaw"], e["fixTime"], e["xWS"], e["uWS"])
        assert ef == 1 and t > 0
        assert obca_b200.ParkingConstraints(e["x0"], e["xF"], e["N"], e["Ts"], e["L"], e["ego"], e["XYbounds"], e["nOb"], e["vOb"], e["A"], e["b"],
                                            xp, up, lp, np_, ts, e["fixTime"], sd) == 1


@pytest.mark.gpu
def test_randomised_sweep_planned_natively_and_solved_in_horizon_groups():
    """main.jl:165-168 as a batch (examples/sweep_parking.py): 48 random start poses planned by libobca_planner.so on host threads,
    grouped by the planner's horizon, solved group by group on the GPU; every converged trajectory passes the reference's checker."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import sweep_parking
    o = sweep_parking.sweep(48, "backwards", 3)
    assert o["planned"] == 48 and o["groups"] >= 5
    assert o["converged"] >= 44 and o["feasible"] >= o["converged"]
