"""Pinning the oracle's IPOPT stand-in with an INDEPENDENT solver: scipy's SLSQP -- an
active-set SQP method that shares nothing with the interior-point code except the NLP callbacks -- is run on the same
reference-formulation NLP (oracle/parking_nlp.py) from the same warm start.  On a small straight-in reverse-parking problem
both must stop at the same primal point; the kernels' own arithmetic (tests/emul) must land there too.
(The OBCA multipliers lambda, mu are not unique where a distance constraint is inactive, so only (x, timeScale, u) and the
objective are compared.)"""
import os
import sys

import numpy as np
import pytest
from scipy.optimize import minimize

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))

import emul                                      # noqa: E402
from obca_b200 import scenarios                  # noqa: E402
from oracle import ipm_ref                       # noqa: E402
from oracle.dualmultws_ref import dualmultws_ipm  # noqa: E402
from oracle.parking_nlp import initial_point     # noqa: E402
from oracle.parking_solve import solve_parking   # noqa: E402


def straight_in(N):
    """Car above the slot of main.jl:99-108, already aligned with it: reverse straight in (feasible in N = 12 steps)."""
    sc = scenarios.reverse_parking_scenario()
    x0 = np.array([0.0, 5.0, np.pi / 2, 0.0])
    xF = sc["xF"]
    Ts, L = 0.6, 2.7
    ys = np.linspace(x0[1], xF[1], N + 1)
    xWS = np.stack([np.zeros(N + 1), ys, np.full(N + 1, np.pi / 2), np.full(N + 1, -(x0[1] - xF[1]) / (N * Ts))], 1)
    xWS[0, 3] = 0.0; xWS[-1, 3] = 0.0
    uWS = np.zeros((N, 2))
    sc.update(B=1, N=N, Ts=Ts, Ts_fix=Ts, L=L, ego=np.array([3.7, 1.0, 1.0, 1.0]), XYbounds=np.array([-15.0, 15.0, 1.0, 10.0]),
              x0=x0[None], rx=xWS[None, :, 0].copy(), ry=xWS[None, :, 1].copy(), ryaw=xWS[None, :, 2].copy(), xWS=xWS[None], uWS=uWS[None])
    return sc


def dense(M):
    return M.toarray() if hasattr(M, "toarray") else np.asarray(M)


@pytest.mark.parametrize("variant,fix", [("d", 0), ("sd", 0), ("sd", 1)])
def test_slsqp_and_ipm_standin_and_kernel_sources_agree(variant, fix):
    N = 12
    sc = straight_in(N)
    a = (sc["x0"][0], sc["xF"], N, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
         sc["rx"][0], sc["ry"][0], sc["ryaw"][0], fix, sc["xWS"][0], sc["uWS"][0])
    out, res, nlp = solve_parking(*a, variant, None, None, ipm_ref.IpmOptions(tol=1e-9, max_iter=300))
    assert res.status == 1
    lay = nlp.lay
    # ---- independent solver: SLSQP on the same callbacks, same starting point as the reference (ParkingSignedDist.jl:213-222) ----
    gL, gU = nlp.gL, nlp.gU
    mL, mU = np.isfinite(gL), np.isfinite(gU)
    cons = [dict(type="eq", fun=nlp.cE, jac=lambda z: dense(nlp.JE(z))),
            dict(type="ineq", fun=lambda z: np.concatenate([(nlp.g(z) - gL)[mL], (gU - nlp.g(z))[mU]]),
                 jac=lambda z: np.vstack([dense(nlp.JI(z))[mL], -dense(nlp.JI(z))[mU]]))]
    lWS, nWS, _, _ = dualmultws_ipm(N, 3, sc["vOb"], sc["A"], sc["b"], sc["rx"][0], sc["ry"][0], sc["ryaw"][0], sc["ego"])
    z_start = initial_point(lay, sc["xWS"][0], sc["uWS"][0], lWS, nWS)
    bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi) for lo, hi in zip(nlp.zL, nlp.zU)]
    r = minimize(nlp.f, z_start, jac=nlp.grad, method="SLSQP", constraints=cons, bounds=bounds, options=dict(ftol=1e-14, maxiter=500))
    assert r.status == 0, r.message
    assert np.abs(nlp.cE(r.x)).max() < 1e-9
    prim = slice(0, lay.oL)                      # x, timeScale, u
    assert np.abs(r.x[prim] - res.z[prim]).max() < 5e-6      # the interior-point iterate sits O(mu / z) inside weakly active bounds
    assert abs(r.fun - nlp.f(res.z)) < 1e-7
    # ---- the kernels' own per-stage source (host build) on the same problem ----
    lp, npp, _, _ = emul.dualmultws_batch(sc)
    o = emul.default_opts(); o.tol = 1e-8; o.mu_min = 1e-9
    k = emul.solve_batch(sc, fix, variant, o, lp, npp)
    assert k["status"][0] == 1
    xp, up, ts = lay.unpack(r.x)[:3]
    tol = 5e-5       # interior point: O(mu / z) inside weakly active bounds
    assert np.abs(k["xp"][0].T - xp).max() < tol and np.abs(k["up"][0].T - up).max() < tol and np.abs(k["ts"][0] - ts).max() < tol


# ----------------------------------------------------------------------------------------------------------------------
# Independent solver with ACTIVE OBCA rows: fixtures of tests/golden/make_slsqp_active.py (SLSQP needs about an hour there)
# ----------------------------------------------------------------------------------------------------------------------
def active_problem(scenario="reverse"):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_slsqp_active as ms
    sc, i = ms.problem(scenario)
    sub = dict(sc); sub.update(B=1, x0=sc["x0"][i:i + 1], rx=sc["rx"][i:i + 1], ry=sc["ry"][i:i + 1], ryaw=sc["ryaw"][i:i + 1],
                               xWS=sc["xWS"][i:i + 1], uWS=sc["uWS"][i:i + 1], Ts_fix=sc["Ts"])
    return sub, ms.N


def active_blocks(sc, N, variant, xp):
    """(obstacle, stages) with the distance row at its bound in the point xp: d == dmin (Dist) / d + sl == dmin with sl at its
    free optimum -0.005 (SignedDist)."""
    from oracle import dualmultws_ref
    g, off = dualmultws_ref.ego_geometry(sc["ego"])
    A = sc["A"]; b = sc["b"].ravel(); vo = np.concatenate([[0], np.cumsum(sc["vOb"])])
    out = []
    for j in range(sc["nOb"]):
        dist = np.array([dualmultws_ref.rect_poly_distance((xp[0, k], xp[1, k], xp[2, k]), A[vo[j]:vo[j + 1]], b[vo[j]:vo[j + 1]], g, off)
                         for k in range(N + 1)])
        act = np.flatnonzero(dist < (0.055 if variant == "sd" else 0.05) + 2e-5)
        if len(act):
            out.append((j, act, slice(vo[j], vo[j + 1])))
    return out


@pytest.mark.parametrize("tag", ["sd", "d_local", "p4_sd_local", "p4_d_local"])
def test_active_rows_slsqp_fixture_vs_interior_point_and_kernel_sources(tag):
    """SLSQP's point (fixture) = the IPOPT stand-in's point (compiled restatement oracle/cpu_ipm, tol 1e-9) = the kernels' own
    source (host build), on a config-2 start pose whose optimum has active distance rows (5 blocks SD / 8 blocks Dist) and on a
    config-3 one (parallel parking, the reference's four obstacles; tags p4_*: 2 / 2 active blocks).
    Primal (x, timeScale, u): 2e-5 (config 3: 5e-5).  lambda, mu ON THE ACTIVE BLOCKS (unique there): 1e-5 / 5e-4 for the kernel sources."""
    from oracle import cpu_ipm
    p4 = tag.startswith("p4_")
    variant = tag.split("_")[1 if p4 else 0]
    fx = np.load(os.path.join(HERE, "golden", "_slsqp", f"slsqp_active_{tag}.npz"))
    sc, N = active_problem("parallel4" if p4 else "reverse")
    assert float(fx["viol"]) < 1e-9
    r = cpu_ipm.ParkingCall(sc, [0], variant, 0).run(opts=cpu_ipm.default_opts(ipm_ref.IpmOptions(tol=1e-9, max_iter=400)),
                                                     lWS=[fx["lWS"]], nWS=[fx["nWS"]])
    assert r["status"][0] == 1
    lay = r["nlp0"].lay
    zs, zi = fx["z"], r["z"][0]
    assert np.abs(zs[:lay.oL] - zi[:lay.oL]).max() < (5e-5 if p4 else 2e-5) and abs(float(fx["f"]) - r["nlp0"].f(zi)) < 1e-7
    xs, us, ts_s, ls, ns = lay.unpack(zs)[:5]
    li, ni = lay.unpack(zi)[3:5]
    blocks = active_blocks(sc, N, variant, xs)
    assert sum(len(a) for _, a, _ in blocks) >= (2 if p4 else 5)
    for j, act, rows in blocks:
        assert np.abs(ls[rows][:, act] - li[rows][:, act]).max() < 1e-5 and np.abs(ns[4 * j:4 * j + 4][:, act] - ni[4 * j:4 * j + 4][:, act]).max() < 1e-5
    # the kernels' per-stage source, host build
    o = emul.default_opts(); o.tol = 1e-8; o.mu_min = 1e-9
    k = emul.solve_batch(sc, 0, variant, o, fx["lWS"][None], fx["nWS"][None])
    assert k["status"][0] == 1
    tolp = 1e-4 if p4 else 5e-5
    assert np.abs(k["xp"][0].T - xs).max() < tolp and np.abs(k["up"][0].T - us).max() < tolp and np.abs(k["ts"][0] - ts_s).max() < tolp
    for j, act, rows in blocks:
        assert np.abs(k["lp"][0].T[rows][:, act] - ls[rows][:, act]).max() < 5e-4
        assert np.abs(k["np"][0].T[4 * j:4 * j + 4][:, act] - ns[4 * j:4 * j + 4][:, act]).max() < 5e-4


def test_config3_far_starts_find_other_local_minima():
    """Recorded finding for config 3 as well: from the far start SLSQP ends in other local minima of both NLPs (objective lower by 2e-3 /
    3e-2, 0.8 m / 2.4 m away) -- fixtures slsqp_active_p4_sd.npz / _d.npz."""
    for variant, gap in (("sd", 1e-3), ("d", 1e-2)):
        far = np.load(os.path.join(HERE, "golden", "_slsqp", f"slsqp_active_p4_{variant}.npz"))
        loc = np.load(os.path.join(HERE, "golden", "_slsqp", f"slsqp_active_p4_{variant}_local.npz"))
        assert float(far["viol"]) < 1e-9 and float(far["f"]) < float(loc["f"]) - gap
        assert np.abs(far["z"][:4 * 21] - loc["z"][:4 * 21]).max() > 0.5


def test_dist_variant_has_a_second_local_minimum():
    """Recorded finding, not a pin: from the far start SLSQP heads for another manoeuvre of the Dist NLP (lower objective, metres
    away from the interior-point solution).  The NLP is non-convex; no stand-in can promise IPOPT's local minimum."""
    from oracle import cpu_ipm
    fx = np.load(os.path.join(HERE, "golden", "_slsqp", "slsqp_active_d.npz"))
    sc, N = active_problem()
    r = cpu_ipm.ParkingCall(sc, [0], "d", 0).run(lWS=[fx["lWS"]], nWS=[fx["nWS"]])
    lay = r["nlp0"].lay
    assert r["status"][0] == 1 and float(fx["f"]) < r["nlp0"].f(r["z"][0]) - 0.05
    assert np.abs(fx["z"][:4 * (N + 1)] - r["z"][0][:4 * (N + 1)]).max() > 1.0


# ----------------------------------------------------------------------------------------------------------
# quadcopter (QuadcopterSignedDist.jl / QuadcopterDist.jl): independent-solver pin with active obstacle rows
# ----------------------------------------------------------------------------------------------------------
def quad_problem():
    """tests/golden/make_slsqp_quad.py: start / goal pair 5 of quadcopter_batch(16, 12, 2), N = 12; four ball-against-box rows active."""
    N, i = 12, 5
    sc = scenarios.quadcopter_batch(16, N, 2)
    return sc, N, i


@pytest.mark.parametrize("variant", ["sd", "d"])
def test_quadcopter_slsqp_fixture_vs_interior_point_and_kernel_sources(variant):
    """SLSQP (active-set SQP; shares only the NLP callbacks) started 1e-3 away from the interior-point solution returns to a KKT point
    (reference-formulation certificate) with the same objective (1e-7 relative) and the same time scale (1e-9); the kernels' host build
    reaches the same objective (3e-5 relative at the default tol = 1e-5) and time scale (1e-5).  The primal trajectory itself is NOT
    pinned by this problem: the objective is flat across the corridor between the boxes (points 0.13 m apart differ by 1e-5 relative in f),
    and from the far start SLSQP finds another route altogether (f = 20.153 against 23.556, fixture *_far) -- the same
    non-convexity finding as for ParkingDist."""
    from oracle import kkt_check
    from oracle.quadcopter_nlp import build_quadcopter_nlp
    from oracle.quadcopter_solve import solve_quadcopter
    sc, N, i = quad_problem()
    g = np.load(os.path.join(HERE, "golden", "_slsqp", f"slsqp_quad_{variant}_local.npz"))
    nlp = build_quadcopter_nlp(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], variant)
    lay = nlp.lay
    e = kkt_check.kkt_certificate(nlp, g["z"])
    assert e["constr_viol"] < 1e-7 and e["E0"] < 5e-5
    gg = nlp.g(g["z"])
    assert int(((gg - nlp.gL) < 1e-6).sum()) == 4                      # four distance rows are active
    out, res, _ = solve_quadcopter(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i], 1.0, variant,
                                   ipm_ref.IpmOptions(tol=1e-9, max_iter=3000), engine="compiled")
    assert res.status == 1
    ts_s = lay.unpack(g["z"])[2]
    assert abs(nlp.f(res.z) - float(g["f"])) < 1e-7 * abs(float(g["f"]))
    assert np.abs(out[2] - ts_s).max() < 1e-9
    sub = {k: (v[i:i + 1] if isinstance(v, np.ndarray) and v.shape[:1] == (16,) else v) for k, v in sc.items()}
    sub["B"] = 1
    r = emul.quad_solve_batch(sub, variant)
    assert r["status"][0] == 1
    zk = lay.pack(r["xp"][0], r["up"][0], r["ts"][0], r["lp"][0], r["slack"][0] if variant == "sd" else None)
    assert abs(nlp.f(zk) - float(g["f"])) < 3e-5 * abs(float(g["f"]))
    assert np.abs(r["ts"][0] - ts_s).max() < 1e-5
    far = np.load(os.path.join(HERE, "golden", "_slsqp", f"slsqp_quad_{variant}_far.npz"))
    assert float(far["f"]) < float(g["f"]) - 1.0 and float(far["viol"]) < 1e-6          # another local minimum, recorded
