"""K1 stand-alone evaluator (obca_b200/csrc/obca_eval.cuh, compiled for the host by tests/emul) against the oracle's
sympy-derived reference NLP: constraint rows, gradient of the Lagrangian, objective -- to round-off."""
import numpy as np
import pytest

import emul
import k1_maps
from obca_b200 import scenarios
from oracle.parking_nlp import build_parking_nlp


@pytest.mark.parametrize("variant,fix", [("sd", 0), ("d", 0), ("sd", 1), ("d", 1)])
def test_k1_matches_oracle(variant, fix):
    sc = scenarios.reverse_parking_scenario()
    N = 14
    rng = np.random.default_rng(5)
    rx, ry, ryaw = rng.normal(size=(3, N + 1))
    x0 = np.array([-6, 9.5, 0.1, 0.2]); xF = sc["xF"]
    nlp = build_parking_nlp(x0, xF, N, 0.7, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"], sc["b"], rx, ry, ryaw, fix, variant)
    z, yE, yI = k1_maps.random_point(nlp, rng)
    arrays, rowmap = k1_maps.k1_inputs(nlp, z, yE, yI)
    one = dict(N=N, nOb=sc["nOb"], vOb=sc["vOb"], A=sc["A"], b=sc["b"], x0=x0, xF=xF, Ts=0.7, L=sc["L"], ego=sc["ego"],
               XYbounds=sc["XYbounds"], rx=rx, ry=ry, ryaw=ryaw)
    c, gl, fk = emul.eval_batch(one, arrays, fix, variant)
    c_ref, gl_ref, f_ref = k1_maps.oracle_reference(nlp, z, yE, yI, rowmap)
    assert np.abs(c - c_ref).max() < 1e-12 * (1 + np.abs(c_ref).max())
    assert np.abs(gl - gl_ref).max() < 1e-11 * (1 + np.abs(gl_ref).max())
    assert abs(fk.sum() - f_ref) < 1e-11 * (1 + abs(f_ref))
