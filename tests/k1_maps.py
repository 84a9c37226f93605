"""Row / variable correspondence between the stand-alone K1 evaluator (obca_eval.cuh, include/obca.h) and the
oracle's restated reference NLP (oracle/parking_nlp.py).  Shared by the CPU (emulation) and GPU tests."""
import numpy as np


def sizes(N, nOb, V, sd):
    NS = N + 1
    n = 4 * NS + NS + 2 * N + V * NS + 4 * nOb * NS + (nOb * NS if sd else 0)
    m = 8 + 6 * N + 4 * nOb * NS
    return n, m


def random_point(nlp, rng):
    """A random (interior-ish) primal point in the oracle's variable order + random multipliers."""
    lay = nlp.lay
    z = rng.normal(size=nlp.n) * 0.3
    if not lay.fixTime:
        z[lay.oT:lay.oU] = 1.0 + 0.1 * rng.normal(size=lay.NS)
    z[lay.oL:lay.oS] = np.abs(z[lay.oL:lay.oS]) + 0.05
    return z, rng.normal(size=nlp.mE), rng.normal(size=nlp.mI)


def k1_inputs(nlp, z, yE, yI):
    """Split the oracle point into the arrays of obca_parking_eval_batch_dev (B = 1) and build y in K1 row order."""
    lay = nlp.lay
    N, NS, nOb, V = lay.N, lay.NS, lay.nOb, lay.V
    sd = lay.variant == "sd"
    xp, up, ts, lp, npp, sl = lay.unpack(z)
    n, m = sizes(N, nOb, V, sd)
    y = np.zeros(m)
    oDyn = 8; oChain = oDyn + 4 * N; oRate = oChain + N; oNorm = oRate + N; oRot = oNorm + nOb * NS; oDist = oRot + 2 * nOb * NS
    rowmap = {}          # (kind, family name) -> K1 row indices in oracle family order
    k = np.arange(N); ks = np.arange(NS)
    rowmap[("eq", "start")] = np.arange(4); rowmap[("eq", "end")] = 4 + np.arange(4)
    for i in range(4):
        rowmap[("eq", f"dyn{i}")] = oDyn + 4 * k + i
    if not lay.fixTime:
        rowmap[("eq", "chain")] = oChain + k
    rowmap[("ineq", "rate0")] = np.array([oRate]); rowmap[("ineq", "rate")] = oRate + np.arange(1, N)
    for j in range(nOb):
        rowmap[("eq" if sd else "ineq", f"norm{j}")] = oNorm + nOb * ks + j
        rowmap[("eq", f"rot1_{j}")] = oRot + 2 * nOb * ks + 2 * j
        rowmap[("eq", f"rot2_{j}")] = oRot + 2 * nOb * ks + 2 * j + 1
        rowmap[("ineq", f"dist{j}")] = oDist + nOb * ks + j
    for fam in nlp.eq:
        y[rowmap[("eq", fam.name)]] = yE[fam.row0:fam.row0 + fam.n]
    for fam in nlp.ineq:
        y[rowmap[("ineq", fam.name)]] = yI[fam.row0:fam.row0 + fam.n]
    arrays = dict(xp=np.ascontiguousarray(xp.T), up=np.ascontiguousarray(up.T), ts=np.ascontiguousarray(ts),
                  lp=np.ascontiguousarray(lp.T), np=np.ascontiguousarray(npp.T),
                  sl=np.ascontiguousarray(sl.T) if sd else None, y=y)
    return arrays, rowmap


def oracle_reference(nlp, z, yE, yI, rowmap):
    """c (K1 row order), gradL (K1 variable order), f from the oracle."""
    lay = nlp.lay
    n, m = sizes(lay.N, lay.nOb, lay.V, lay.variant == "sd")
    c = np.zeros(m)
    cE = nlp.cE(z); g = nlp.g(z)
    for fam in nlp.eq:
        c[rowmap[("eq", fam.name)]] = cE[fam.row0:fam.row0 + fam.n]
    for fam in nlp.ineq:
        c[rowmap[("ineq", fam.name)]] = g[fam.row0:fam.row0 + fam.n]
    gl = nlp.grad(z) + nlp.JE(z).T @ yE + nlp.JI(z).T @ yI
    if lay.fixTime:      # K1 keeps an (all-zero) timeScale block
        gl = np.concatenate([gl[:lay.oT], np.zeros(lay.NS), gl[lay.oT:]])
    return c, gl, nlp.f(z)
