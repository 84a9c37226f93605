"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of the reference quadcopter NLPs in the reference's own formulation:

  QuadcopterSignedDist -> QuadcopterNavigation/QuadcopterSignedDist.jl:25-300  (variant="sd")
  QuadcopterDist       -> QuadcopterNavigation/QuadcopterDist.jl:25-282        (variant="d")

including quirk Q5 of SURVEY.md A.4: `x[11]*x[12]` etc. in the body-rate rows (QuadcopterSignedDist.jl:153-155) is
LINEAR indexing into the 12x(N+1) array, i.e. always the stage-1 entries x[11,1]*x[12,1] (pinned to x0 by :108-119).
Parity status: UNPINNED (no golden vectors in the reference, Julia/Ipopt not installed).

Variable layout (0-based, JuMP declaration order :34-49):
  x[i,k] = 12k+i | ts[k] = oT+k | u[i,k] = oU+4k+i | l_o[r,k] = oL + o*6*NS + 6k + r (o = obstacle 0..4) |
  slack[j,k] = oS + 5k + j  ("sd" only)
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .sparse_nlp import Family, SparseNLP, Template

MASS, GRAV, K_F, K_M, ARM = 0.5, 9.81, 0.0611, 0.0015, 0.225          # :51-62
INERT = (3.9e-3, 4.4e-3, 4.9e-3)
W_H = float(np.sqrt(MASS * GRAV / (K_F * 4)))
REG2, REG3 = 1e-4, 1e-4


class QLayout:
    def __init__(self, N, variant):
        self.N = int(N); self.NS = N + 1; self.variant = variant
        NS = self.NS
        self.oX = 0; self.oT = 12 * NS; self.oU = self.oT + NS; self.oL = self.oU + 4 * N
        self.oS = self.oL + 30 * NS
        self.n = self.oS + (5 * NS if variant == "sd" else 0)

    def x(self, i, k): return 12 * np.asarray(k) + i
    def ts(self, k): return self.oT + np.asarray(k)
    def u(self, i, k): return self.oU + 4 * np.asarray(k) + i
    def l(self, o, r, k): return self.oL + o * 6 * self.NS + 6 * np.asarray(k) + r
    def sl(self, j, k): return self.oS + 5 * np.asarray(k) + j

    def unpack(self, z):
        NS, N = self.NS, self.N
        xp = z[:12 * NS].reshape(NS, 12).T.copy()
        ts = z[self.oT:self.oT + NS].copy()
        up = z[self.oU:self.oU + 4 * N].reshape(N, 4).T.copy()
        lp = np.concatenate([z[self.oL + o * 6 * NS:self.oL + (o + 1) * 6 * NS].reshape(NS, 6).T for o in range(5)], 0)   # 30 x NS (:296)
        sl = z[self.oS:].reshape(NS, 5).T.copy() if self.variant == "sd" else None
        return xp, up, ts, lp, sl

    def pack(self, xp, up, ts, lp, sl=None):
        z = np.zeros(self.n); NS = self.NS
        z[:12 * NS] = np.asarray(xp, float).T.ravel()
        z[self.oT:self.oT + NS] = np.asarray(ts, float).ravel()
        z[self.oU:self.oU + 4 * self.N] = np.asarray(up, float).T.ravel()
        lp = np.asarray(lp, float)
        for o in range(5):
            z[self.oL + o * 6 * NS:self.oL + (o + 1) * 6 * NS] = lp[6 * o:6 * o + 6].T.ravel()
        if self.variant == "sd":
            z[self.oS:] = np.asarray(sl, float).T.ravel()
        return z


_T: dict = {}


def _dyn_templates():
    """QuadcopterSignedDist.jl:136-158.  vars: x[0..11] (stage i), u[0..3], ts, xn[0..11] (stage i+1), and the three
    stage-1 body rates p1,q1,r1 of quirk Q5; param Ts."""
    x = sp.symbols("x0:12"); u = sp.symbols("u0:4"); ts, Ts = sp.symbols("ts Ts"); xn = sp.symbols("xn0:12")
    p1, q1, r1 = sp.symbols("p1 q1 r1")        # x[10,1], x[11,1], x[12,1]
    h = ts * Ts
    F = K_F * sum(ui ** 2 for ui in u)
    f = [
        x[6], x[7], x[8],
        sp.cos(x[4]) * x[9] + sp.sin(x[4]) * x[11],
        sp.sin(x[4]) * sp.tan(x[3]) * x[9] + x[10] - sp.cos(x[4]) * sp.tan(x[3]) * x[11],
        -sp.sin(x[4]) * sp.sec(x[3]) * x[9] + sp.cos(x[4]) * sp.sec(x[3]) * x[11],
        1 / MASS * (F * (sp.sin(x[3]) * sp.cos(x[4]) * sp.sin(x[5]) + sp.sin(x[4]) * sp.cos(x[5]))),
        1 / MASS * (F * (-sp.sin(x[3]) * sp.cos(x[4]) * sp.cos(x[5]) + sp.sin(x[4]) * sp.sin(x[5]))),
        1 / MASS * (F * (sp.cos(x[3]) * sp.cos(x[4])) - MASS * GRAV),
        1 / INERT[0] * (ARM * K_F * (u[1] ** 2 - u[3] ** 2) - (INERT[2] - INERT[1]) * q1 * r1),
        1 / INERT[1] * (ARM * K_F * (u[2] ** 2 - u[0] ** 2) - (INERT[0] - INERT[2]) * p1 * r1),
        1 / INERT[2] * (K_M * (u[0] ** 2 - u[1] ** 2 + u[2] ** 2 - u[3] ** 2) - (INERT[1] - INERT[0]) * p1 * q1),
    ]
    vs = list(x) + list(u) + [ts] + list(xn) + [p1, q1, r1]
    return [Template(xn[i] - (x[i] + h * f[i]), vs, [Ts]) for i in range(12)]


def build_quadcopter_nlp(x0, xF, N, Ts, R, obs, variant="sd"):
    """obs: 5 box vectors b (6 each) = ob1..ob5 of QuadcopterSignedDist.jl:25 (A = [I; -I], :162-163)."""
    x0 = np.asarray(x0, float).ravel(); xF = np.asarray(xF, float).ravel()
    obs = [np.asarray(o, float).ravel() for o in obs]
    lay = QLayout(N, variant)
    NS = lay.NS
    nlp = SparseNLP(lay.n); nlp.lay = lay
    ks = np.arange(NS); kN = np.arange(N)
    u, un, q, ts, c0, sl = sp.symbols("u un q ts c0 sl")
    # ---- objective :65-71 ----
    t = _T.setdefault("hover", Template(1e-3 * (W_H - u) ** 2, [u], []))
    jj, kk = np.meshgrid(np.arange(4), kN, indexing="ij")
    nlp.add_obj(Family("hover", t, lay.u(jj.ravel(), kk.ravel()).reshape(-1, 1)))
    t = _T.setdefault("du", Template(1e-2 * (u - un) ** 2, [u, un], []))
    jj, kk = np.meshgrid(np.arange(4), np.arange(N - 1), indexing="ij")
    nlp.add_obj(Family("du", t, np.stack([lay.u(jj.ravel(), kk.ravel()), lay.u(jj.ravel(), kk.ravel() + 1)], 1)))
    t = _T.setdefault("rate", Template(REG3 * q ** 2, [q], []))
    jj, kk = np.meshgrid(np.array([9, 10, 11]), ks, indexing="ij")
    nlp.add_obj(Family("rates", t, lay.x(jj.ravel(), kk.ravel()).reshape(-1, 1)))
    t = _T.setdefault("ts", Template(0.25 * ts + 5 * ts ** 2, [ts], []))
    nlp.add_obj(Family("ts", t, lay.ts(ks).reshape(-1, 1)))
    if variant == "sd":
        t = _T.setdefault("slack", Template(1e2 * sl + 1e3 * sl ** 2, [sl], []))
        jj, kk = np.meshgrid(np.arange(5), ks, indexing="ij")
        nlp.add_obj(Family("slack", t, lay.sl(jj.ravel(), kk.ravel()).reshape(-1, 1)))
    t = _T.setdefault("lreg", Template(REG2 * q ** 2, [q], []))
    nlp.add_obj(Family("lreg", t, np.arange(lay.oL, lay.oS).reshape(-1, 1)))
    # ---- bounds :74-105 ----
    for j in range(4):
        nlp.zL[lay.u(j, kN)] = 1.2; nlp.zU[lay.u(j, kN)] = 7.8
    lo = [0, 0, 0, -3, -0.2, -0.2, -1, -1, -1, -1 if variant == "sd" else -1.5, -1, -1]
    hi = [10, 10, 5, 3, 0.2, 0.2, 1, 1, 1, 1 if variant == "sd" else 3, 1, 1]
    for i in range(12):
        nlp.zL[lay.x(i, ks)] = lo[i]; nlp.zU[lay.x(i, ks)] = hi[i]
    nlp.zL[lay.ts(ks)] = 0.5; nlp.zU[lay.ts(ks)] = 2.0
    nlp.zL[lay.oL:lay.oS] = 0.0
    if variant == "sd":
        nlp.zL[lay.oS:] = 0.0
    # ---- start / end :108-134 ----
    t = _T.setdefault("fix", Template(q - c0, [q], [c0]))
    nlp.add_eq(Family("start", t, lay.x(np.arange(12), 0).reshape(-1, 1), x0.reshape(-1, 1)))
    nlp.add_eq(Family("end", t, lay.x(np.arange(12), N).reshape(-1, 1), xF.reshape(-1, 1)))
    # ---- dynamics + chain :136-158 ----
    if "dyn" not in _T:
        _T["dyn"] = _dyn_templates()
    cols = [lay.x(i, kN) for i in range(12)] + [lay.u(j, kN) for j in range(4)] + [lay.ts(kN)] + \
           [lay.x(i, kN + 1) for i in range(12)] + [np.full(N, lay.x(9, 0)), np.full(N, lay.x(10, 0)), np.full(N, lay.x(11, 0))]
    idx = np.stack(cols, 1)
    for i in range(12):
        nlp.add_eq(Family(f"dyn{i}", _T["dyn"][i], idx, np.full((N, 1), Ts)))
    t1, t2 = sp.symbols("t1 t2")
    t = _T.setdefault("chain", Template(t1 - t2, [t1, t2], []))
    nlp.add_eq(Family("chain", t, np.stack([lay.ts(kN), lay.ts(kN + 1)], 1)))
    # ---- OBCA rows :165-197 ----
    if "obca" not in _T:
        X, Y, Z = sp.symbols("X Y Z"); lam = sp.symbols("lam0:6"); bb = sp.symbols("bb0:6")
        p = [lam[0] - lam[3], lam[1] - lam[4], lam[2] - lam[5]]
        norm = p[0] ** 2 + p[1] ** 2 + p[2] ** 2 - 1
        dist = -sum(bb[r] * lam[r] for r in range(6)) + X * p[0] + Y * p[1] + Z * p[2]
        _T["obca"] = (Template(norm, list(lam), []), Template(dist + 0.01 * sl, [X, Y, Z] + list(lam) + [sl], list(bb)),
                      Template(dist, [X, Y, Z] + list(lam), list(bb)))
    tn, tds, tdd = _T["obca"]
    for o in range(5):
        lidx = np.stack([lay.l(o, r, ks) for r in range(6)], 1)
        nlp.add_eq(Family(f"norm{o}", tn, lidx))                       # "==" in BOTH variants (QuadcopterDist.jl:163)
        pos = np.stack([lay.x(0, ks), lay.x(1, ks), lay.x(2, ks)], 1)
        par = np.tile(obs[o], (NS, 1))
        if variant == "sd":
            nlp.add_ineq(Family(f"dist{o}", tds, np.concatenate([pos, lidx, lay.sl(o, ks).reshape(-1, 1)], 1), par, lo=R))
        else:
            nlp.add_ineq(Family(f"dist{o}", tdd, np.concatenate([pos, lidx], 1), par, lo=R))
    return nlp


def dual_ws(pos, b6):
    """Closed-form dual warm start of one (position, box) pair (twin of quad_dual_ws in obca_quad_local.cuh)."""
    hi = b6[:3]; lo = -b6[3:]
    cl = np.clip(pos, lo, hi)
    d = pos - cl
    lam = np.zeros(6)
    n2 = d @ d
    if n2 > 1e-24:
        p = d / np.sqrt(n2)
        lam[:3] = np.maximum(p, 0); lam[3:] = np.maximum(-p, 0)
    else:
        pen = np.concatenate([hi - pos, pos - lo])
        # same tie-breaking order as the kernel: (hi_0, lo_0, hi_1, lo_1, hi_2, lo_2), first strict minimum
        order = [0, 3, 1, 4, 2, 5]
        best = min(order, key=lambda i: (pen[i], order.index(i)))
        lam[best] = 1.0
    return lam


def initial_point(lay: QLayout, xWS, timeWS, obs=None):
    """:199-210: ts = timeWS, x = xWS, u = w_H ("faster not to warm-start", uWS ignored), slack = 1 and
    l = 0.05 (obs=None, the reference's start) or the closed-form dual warm start (obs given)."""
    z = np.zeros(lay.n)
    xWS = np.asarray(xWS, float)
    z[:12 * lay.NS] = xWS[:, :lay.NS].T.ravel()
    z[lay.oT:lay.oT + lay.NS] = timeWS
    z[lay.oU:lay.oU + 4 * lay.N] = W_H
    z[lay.oL:lay.oS] = 0.05
    if obs is not None:
        for o in range(5):
            for k in range(lay.NS):
                z[lay.l(o, np.arange(6), k)] = dual_ws(xWS[:3, k], np.asarray(obs[o], float))
    if lay.variant == "sd":
        z[lay.oS:] = 1.0
    return z


def stage_order(nlp):
    lay = nlp.lay; NS = lay.NS
    vs = np.zeros(nlp.n, int)
    vs[:12 * NS] = np.repeat(np.arange(NS), 12)
    vs[lay.oT:lay.oT + NS] = np.arange(NS)
    vs[lay.oU:lay.oU + 4 * lay.N] = np.repeat(np.arange(lay.N), 4)
    for o in range(5):
        vs[lay.oL + o * 6 * NS:lay.oL + (o + 1) * 6 * NS] = np.repeat(np.arange(NS), 6)
    if lay.variant == "sd":
        vs[lay.oS:] = np.repeat(np.arange(NS), 5)
    rs = np.zeros(nlp.mE, int)
    for fam in nlp.eq:
        rs[fam.row0:fam.row0 + fam.n] = vs[fam.idx].max(axis=1)
    return np.argsort(np.concatenate([2 * vs, 2 * rs + 1]), kind="stable")
