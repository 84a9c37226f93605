"""Quadcopter NLPs (QuadcopterSignedDist.jl / QuadcopterDist.jl / constrSatisfaction.jl, BASELINE config 4).
CPU part: kernel sources (host emulation) against the oracle.  GPU part (-m gpu): the C-ABI path."""
import numpy as np
import pytest

from obca_b200 import scenarios
from oracle import checkers, ipm_ref, kkt_check
from oracle.quadcopter_nlp import build_quadcopter_nlp
from oracle.quadcopter_solve import solve_quadcopter


def _cert(sc, i, N, variant, xp, up, ts, lp, sl):
    nlp = build_quadcopter_nlp(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], variant)
    z = nlp.lay.pack(xp, up, ts, lp, sl if variant == "sd" else None)
    return kkt_check.kkt_certificate(nlp, z), nlp


def test_oracle_quadcopter_nlp_sizes_and_derivatives():
    sc = scenarios.quadcopter_batch(1, 10, 2)
    for variant, n_exp in (("sd", 52 * 10 + 48), ("d", 47 * 10 + 43)):
        nlp = build_quadcopter_nlp(sc["x0"][0], sc["xF"][0], 10, sc["Ts"], sc["R"], sc["obs"], variant)
        assert (nlp.n, nlp.mE, nlp.mI) == (n_exp, 18 * 10 + 29, 5 * 10 + 5)
        rng = np.random.default_rng(0)
        z = rng.normal(size=nlp.n) * 0.2
        z[nlp.lay.oT:nlp.lay.oU] = 1.0; z[nlp.lay.oU:nlp.lay.oL] += 4.0
        yE = rng.normal(size=nlp.mE); yI = rng.normal(size=nlp.mI)
        H = nlp.hess(z, yE, yI).toarray(); JE = nlp.JE(z).toarray()
        gl = lambda zz: nlp.grad(zz) + nlp.JE(zz).T @ yE + nlp.JI(zz).T @ yI
        for j in rng.choice(nlp.n, 40, replace=False):
            e = np.zeros(nlp.n); e[j] = 1e-6
            assert np.abs((nlp.cE(z + e) - nlp.cE(z - e)) / 2e-6 - JE[:, j]).max() < 1e-6
            assert np.abs((gl(z + e) - gl(z - e)) / 2e-6 - H[:, j]).max() < 2e-5


@pytest.mark.parametrize("variant", ["sd", "d"])
def test_kernel_sources_vs_oracle(variant):
    import emul
    N = 20
    sc = scenarios.quadcopter_batch(2, N, 2)
    i = 0
    # first Newton step: identical up to round-off x the 1/dc terminal penalty
    o = emul.default_opts(); o.max_iter = 1; o.dc = 1e-6
    r = emul.quad_solve_batch(sc, variant, o)
    out, res, nlp = solve_quadcopter(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i], 1.0, variant,
                                     ipm_ref.IpmOptions(max_iter=1, dc_value=1e-6))
    assert np.abs(out[0] - r["xp"][i]).max() < 1e-7 and np.abs(out[1] - r["up"][i]).max() < 1e-7
    assert np.abs(out[2] - r["ts"][i]).max() < 1e-8 and np.abs(out[5] - r["lp"][i]).max() < 1e-7
    # converged: same objective and time scale, both KKT points of the reference NLP, verbatim checker passes
    o = emul.default_opts(); o.max_iter = 3000
    r = emul.quad_solve_batch(sc, variant, o)
    out, res, nlp = solve_quadcopter(sc["x0"][i], sc["xF"][i], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i], 1.0, variant,
                                     ipm_ref.IpmOptions(max_iter=3000, linsolve="sparse"))
    assert r["status"][i] == 1 and res.status == 1
    e, nlp2 = _cert(sc, i, N, variant, r["xp"][i], r["up"][i], r["ts"][i], r["lp"][i], r["slack"][i])
    assert e["E0"] < 5e-5
    assert abs(e["f"] - nlp.f(res.z)) < 1e-4 * abs(e["f"]) and np.abs(out[2] - r["ts"][i]).max() < 1e-5
    assert checkers.constrSatisfaction(r["xp"][i], r["up"][i], r["ts"][i], sc["x0"][i], sc["xF"][i], sc["Ts"], r["lp"][i], *sc["obs"], sc["R"])
    if variant == "sd":
        assert r["slack"][i].sum() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["sd", "d"])
def test_gpu_quadcopter_config4(variant):
    import obca_b200
    from obca_b200 import quadcopter
    N, B = 100, 16
    sc = scenarios.quadcopter_batch(B, N, 2)
    r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1 if variant == "sd" else 0)
    assert (r["exitflag"] == 1).sum() >= B - 1
    feas, worst = quadcopter.check_quadcopter_batch(r["xp"], r["up"], r["ts"], sc["x0"], sc["xF"], sc["Ts"], r["lp"], sc["obs"], sc["R"])
    for i in range(B):
        ref = checkers.constrSatisfaction(r["xp"][i], r["up"][i], r["ts"][i], sc["x0"][i], sc["xF"][i], sc["Ts"], r["lp"][i], *sc["obs"], sc["R"])
        assert bool(feas[i]) == ref
        if r["exitflag"][i] == 1:
            assert ref
    i = int(np.argmax(r["exitflag"] == 1))
    e, _ = _cert(sc, i, N, variant, r["xp"][i], r["up"][i], r["ts"][i], r["lp"][i], r["slack"][i])
    assert e["E0"] < 1e-4 and e["constr_viol"] < 1e-4
    # reference-named single-problem call surface (QuadcopterSignedDist.jl:25 / :298)
    f = obca_b200.QuadcopterSignedDist if variant == "sd" else obca_b200.QuadcopterDist
    xp, up, tsp, ef, t, lp, status = f(sc["x0"][i][None], sc["xF"][i][None], N, sc["Ts"], sc["R"], *sc["obs"], sc["xWS"][i], np.full((4, N), 0.5), 1)
    assert xp.shape == (12, N + 1) and up.shape == (4, N) and tsp.shape == (N + 1,) and lp.shape == (30, N + 1) and ef == 1 and status == "Optimal"
    assert obca_b200.constrSatisfaction(xp, up, tsp, sc["x0"][i][None], sc["xF"][i][None], sc["Ts"], lp, *sc["obs"], sc["R"]) is True


@pytest.mark.gpu
@pytest.mark.parametrize("N", [20, 40, 70, 100])
def test_gpu_tensor_core_sweep_equals_host_recursion(N):
    """The device sweep (dense zero-padded tiles, DMMA products, kkt_solve_block) against the plain one-thread recursion of the same
    elimination (kkt_dense, host build of the same sources) -- one, two, three and four warps per problem (N + 1 rounded up to 32
    threads): the first Newton step and the iterate after five iterations agree up to round-off x the terminal penalty 1/dc (as in
    test_kernel_sources_vs_oracle, dc = 1e-6 for this comparison); then the full solve converges to a KKT point that passes the verbatim checker."""
    import emul
    import obca_b200
    from obca_b200 import quadcopter
    B = 3
    sc = scenarios.quadcopter_batch(B, N, 2)
    for variant in ("sd", "d"):
        for it, tol in ((1, 1e-8), (5, 1e-4)):
            oe = emul.default_opts(); oe.max_iter = it; oe.dc = 1e-6
            og = obca_b200.default_opts(); og.max_iter = it; og.dc = 1e-6
            re_ = emul.quad_solve_batch(sc, variant, oe)
            rg = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1 if variant == "sd" else 0, og)
            for k in ("xp", "up", "ts", "lp"):
                scale = 1.0 + np.abs(re_[k]).max()
                assert np.abs(rg[k] - re_[k]).max() < tol * scale, (variant, it, k, np.abs(rg[k] - re_[k]).max())
        r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1 if variant == "sd" else 0)
        assert (r["exitflag"] >= 1).all()
        feas, _ = quadcopter.check_quadcopter_batch(r["xp"], r["up"], r["ts"], sc["x0"], sc["xF"], sc["Ts"], r["lp"], sc["obs"], sc["R"])
        assert feas[r["exitflag"] == 1].all()
        i = int(np.argmax(r["exitflag"] == 1))
        e, _ = _cert(sc, i, N, variant, r["xp"][i], r["up"][i], r["ts"][i], r["lp"][i], r["slack"][i])
        assert e["E0"] < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["quad_sd_p0", "quad_d_p1"])
def test_gpu_quadcopter_matches_oracle_golden(case):
    """BASELINE config 4 (N = 100): the library's solution against the golden of the IPOPT stand-in on the restated reference NLP
    (tests/golden/make_golden.py), same inputs.  Primal trajectory 5e-4 (two interior-point runs stopped at tol 1e-5; the
    quadcopter objective is flat along the time-optimal path), time scale 1e-4, objective 1e-4 relative."""
    import os
    from obca_b200 import quadcopter
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    N = int(g["N"]); i = int(g["index"]); variant = str(g["variant"])
    sc = scenarios.quadcopter_batch(4, N, int(g["seed"]))
    assert int(g["status"]) == 1
    r = quadcopter.quadcopter_solve_batch(sc["x0"][i:i + 1], sc["xF"][i:i + 1], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i:i + 1], 1.0,
                                          1 if variant == "sd" else 0)
    assert r["exitflag"][0] == int(g["exitflag"])
    assert np.abs(r["ts"][0] - g["ts"]).max() < 1e-4
    assert np.abs(r["xp"][0] - g["xp"]).max() < 5e-4 and np.abs(r["up"][0] - g["up"]).max() < 5e-4
    e, _ = _cert(sc, i, N, variant, r["xp"][0], r["up"][0], r["ts"][0], r["lp"][0], r["slack"][0])
    assert e["E0"] < 1e-4 and abs(e["f"] - float(g["f"])) < 1e-4 * abs(float(g["f"]))


@pytest.mark.gpu
def test_gpu_quadcopter_full_config4_batch():
    """BASELINE config 4 at its full size (B = 2048, N = 100, SD): convergence rate, the verbatim constrSatisfaction on every
    returned trajectory (GPU twin), the slack gate of QuadcopterSignedDist.jl:283-288, and the reference-formulation KKT
    certificate on a sample."""
    from obca_b200 import quadcopter
    N, B = 100, 2048
    sc = scenarios.quadcopter_batch(B, N, 2)
    r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1)
    ok = r["exitflag"] >= 1
    assert ok.mean() >= 0.97, ok.mean()
    feas, worst = quadcopter.check_quadcopter_batch(r["xp"], r["up"], r["ts"], sc["x0"], sc["xF"], sc["Ts"], r["lp"], sc["obs"], sc["R"])
    assert feas[r["exitflag"] == 1].all()
    gate = r["slack"].reshape(B, -1).sum(1) > 1e-3
    assert np.array_equal(r["exitflag"][ok] == 2, gate[ok])
    for i in np.flatnonzero(r["exitflag"] == 1)[:3]:
        e, _ = _cert(sc, int(i), N, "sd", r["xp"][i], r["up"][i], r["ts"][i], r["lp"][i], r["slack"][i])
        assert e["E0"] < 1e-4 and e["constr_viol"] < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["sd", "d"])
def test_gpu_quadcopter_objective_equals_independent_sqp(variant):
    """The SLSQP fixture of tests/golden/make_slsqp_quad.py (N = 12, four active ball-against-box rows; an active-set SQP solver that
    shares only the NLP callbacks with the oracle): the library reaches the same objective (3e-5 relative at tol = 1e-5) and the same
    time scale (1e-5), and its point is a KKT point of the reference NLP.  (The trajectory itself is not pinned by this problem -- the
    objective is flat across the corridor between the boxes, tests/test_oracle_cross_solver.py.)"""
    import os
    from obca_b200 import quadcopter
    N, i = 12, 5
    sc = scenarios.quadcopter_batch(16, N, 2)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "_slsqp", f"slsqp_quad_{variant}_local.npz"))
    r = quadcopter.quadcopter_solve_batch(sc["x0"][i:i + 1], sc["xF"][i:i + 1], N, sc["Ts"], sc["R"], sc["obs"], sc["xWS"][i:i + 1], 1.0,
                                          1 if variant == "sd" else 0)
    assert r["exitflag"][0] == 1
    e, nlp = _cert(sc, i, N, variant, r["xp"][0], r["up"][0], r["ts"][0], r["lp"][0], r["slack"][0])
    assert e["E0"] < 1e-4 and e["constr_viol"] < 1e-4
    assert abs(e["f"] - float(g["f"])) < 3e-5 * abs(float(g["f"]))
    assert np.abs(r["ts"][0] - nlp.lay.unpack(g["z"])[2]).max() < 1e-5
