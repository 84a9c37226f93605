"""Oracle self-checks (CPU): the restated reference NLP has the sizes the reference model has, its sympy
derivatives agree with finite differences, and the input producers reproduce the known answers of SURVEY.md 8c."""
import numpy as np
import pytest

from obca_b200 import scenarios
from oracle import dualmultws_ref
from oracle.parking_nlp import build_parking_nlp


def _nlp(variant, fix, N=12):
    sc = scenarios.reverse_parking_scenario()
    rng = np.random.default_rng(1)
    rx, ry, ryaw = rng.normal(size=(3, N + 1))
    return build_parking_nlp([-6, 9.5, 0, 0], sc["xF"], N, 0.6, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"],
                             sc["A"], sc["b"], rx, ry, ryaw, fix, variant)


def test_obst_hrep_known_answer():
    sc = scenarios.reverse_parking_scenario()     # main.jl:102-104 through obstHrep.jl:57-72
    assert np.array_equal(sc["A"], np.array([[0, 1], [1, 0], [-1, 0], [0, 1], [0, -1.0]]))
    assert np.allclose(sc["b"].ravel(), [5, -1.3, -1.3, 5, -11])
    assert list(sc["vOb"]) == [2, 2, 1]
    # general (slanted) edge, obstHrep.jl:73-85: the row is NOT normalised
    A, b = scenarios.obst_hrep(1, [3], [[[0, 0], [1, 2], [3, 1]]])
    assert np.allclose(A, [[-2, 1], [0.5, 1]]) and np.allclose(b.ravel(), [0, 2.5])


@pytest.mark.parametrize("variant,fix,N", [("sd", 0, 80), ("d", 0, 80), ("sd", 1, 40)])
def test_sizes_match_survey_A6(variant, fix, N):
    nlp = _nlp(variant, fix, N)
    if variant == "sd" and not fix:
        assert (nlp.n, nlp.mE, nlp.mI, nlp.nnz_jac()) == (27 * N + 25, 14 * N + 17, 4 * N + 3, 91 * N + 69)
    elif variant == "d" and not fix:
        assert (nlp.n, nlp.mE, nlp.mI, nlp.nnz_jac()) == (24 * N + 22, 11 * N + 14, 7 * N + 6, 88 * N + 66)
    else:
        assert nlp.n == 26 * N + 24 and nlp.mE == 13 * N + 17


@pytest.mark.parametrize("variant,fix", [("sd", 0), ("d", 0), ("sd", 1), ("d", 1)])
def test_derivatives_vs_finite_differences(variant, fix):
    nlp = _nlp(variant, fix)
    rng = np.random.default_rng(0)
    z = rng.normal(size=nlp.n) * 0.3
    if not fix:
        z[nlp.lay.oT:nlp.lay.oU] = 1.0 + 0.1 * rng.normal(size=nlp.lay.NS)
    yE = rng.normal(size=nlp.mE); yI = rng.normal(size=nlp.mI)
    g = nlp.grad(z); JE = nlp.JE(z).toarray(); JI = nlp.JI(z).toarray(); H = nlp.hess(z, yE, yI).toarray()
    h = 1e-6
    idx = rng.choice(nlp.n, 60, replace=False)
    for i in idx:
        e = np.zeros(nlp.n); e[i] = h
        assert abs((nlp.f(z + e) - nlp.f(z - e)) / (2 * h) - g[i]) < 1e-5 * (1 + abs(g[i]))
        assert np.abs((nlp.cE(z + e) - nlp.cE(z - e)) / (2 * h) - JE[:, i]).max() < 1e-6
        assert np.abs((nlp.g(z + e) - nlp.g(z - e)) / (2 * h) - JI[:, i]).max() < 1e-6
        gl = lambda zz: nlp.grad(zz) + nlp.JE(zz).T @ yE + nlp.JI(zz).T @ yI
        assert np.abs((gl(z + e) - gl(z - e)) / (2 * h) - H[:, i]).max() < 1e-5
    assert np.abs(H - H.T).max() == 0


def test_dualmultws_known_answers():
    """SURVEY.md 8c known-answer vectors: optimum d* = rectangle/polyhedron distance."""
    sc = scenarios.reverse_parking_scenario()
    g, off = dualmultws_ref.ego_geometry(sc["ego"])
    A, b = sc["A"], sc["b"].ravel()
    for pose, rows, exp in [((-6, 9.5, 0), [4], 0.5), ((-6, 9.5, 0), [0, 1], 3.5), ((0, 1.3, np.pi / 2), [0, 1], 0.3),
                            ((5, 9.5, 0.3), [2, 3], 3.2491433042)]:
        lam, mu, d = dualmultws_ref.solve_one(A[rows], b[rows], pose, g, off)
        assert abs(d - exp) < 1e-6
        assert abs(dualmultws_ref.rect_poly_distance(pose, A[rows], b[rows], g, off) - exp) < 1e-8
    lam, mu, d = dualmultws_ref.solve_one(A[[2, 3]], b[[2, 3]], (5, 9.5, 0.3), g, off)
    assert np.allclose(lam, [0, 1], atol=1e-6) and np.allclose(mu, [0, 0, 0.2955202, 0.9553365], atol=1e-5)
    # overlapping pose: optimum 0 with lam = mu = 0
    lam, mu, d = dualmultws_ref.solve_one(A[[0, 1]], b[[0, 1]], (-3.0, 4.0, 0.2), g, off)
    assert abs(d) < 1e-7
