"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of the reference parking NLP, line by line, in the reference's OWN
formulation (JuMP variable order, one expression per @NLconstraint, bounds kept
as two-sided rows on single variables):

  ParkingSignedDist  -> AutonomousParking/ParkingSignedDist.jl:29-314  (variant="sd")
  ParkingDist        -> AutonomousParking/ParkingDist.jl:29-315        (variant="d")

Parity status: UNPINNED -- the reference ships no golden vectors and Julia /
JuMP / Ipopt are not available in this image; the
restatement is validated by finite differences (tests/test_oracle_nlp.py) and
by two independent solvers agreeing on its KKT points.

Variable layout (0-based, JuMP declaration order ParkingSignedDist.jl:49-59):
  x[i,k]  = 4*k+i                    k=0..N     (X, Y, psi, v)
  ts[k]   = oT+k                     k=0..N     (only if fixTime==0)
  u[i,k]  = oU+2*k+i                 k=0..N-1   (delta, a)
  l[r,k]  = oL+V*k+r                 k=0..N     V=sum(vOb)
  n[r,k]  = oN+4*nOb*k+r             k=0..N
  sl[j,k] = oS+nOb*k+j               k=0..N     ("sd" only)
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .sparse_nlp import Family, SparseNLP, Template

DMIN = 0.05  # ParkingSignedDist.jl:33


class Layout:
    def __init__(self, N, nOb, vOb, fixTime, variant):
        self.N, self.nOb = int(N), int(nOb)
        self.vOb = [int(v) for v in np.asarray(vOb).ravel()]
        self.V = sum(self.vOb)
        self.fixTime = int(fixTime)
        self.variant = variant
        NS = N + 1
        self.NS = NS
        self.oX = 0
        self.oT = 4 * NS
        self.oU = self.oT + (0 if fixTime else NS)
        self.oL = self.oU + 2 * N
        self.oN = self.oL + self.V * NS
        self.oS = self.oN + 4 * nOb * NS
        self.n = self.oS + (nOb * NS if variant == "sd" else 0)
        self.voff = np.concatenate([[0], np.cumsum(self.vOb)]).astype(int)

    def x(self, i, k): return 4 * np.asarray(k) + i
    def ts(self, k): return self.oT + np.asarray(k)
    def u(self, i, k): return self.oU + 2 * np.asarray(k) + i
    def l(self, r, k): return self.oL + self.V * np.asarray(k) + r
    def mu(self, r, k): return self.oN + 4 * self.nOb * np.asarray(k) + r
    def sl(self, j, k): return self.oS + self.nOb * np.asarray(k) + j

    # pack / unpack helpers (column-major Julia shapes)
    def pack(self, xp, up, ts, lp, np_, sl=None):
        z = np.zeros(self.n)
        z[self.oX:self.oX + 4 * self.NS] = np.asarray(xp, float).reshape(4, self.NS).T.ravel()
        if not self.fixTime:
            z[self.oT:self.oT + self.NS] = np.asarray(ts).ravel()
        z[self.oU:self.oU + 2 * self.N] = np.asarray(up).T.ravel()
        z[self.oL:self.oL + self.V * self.NS] = np.asarray(lp).T.ravel()
        z[self.oN:self.oN + 4 * self.nOb * self.NS] = np.asarray(np_).T.ravel()
        if self.variant == "sd":
            z[self.oS:] = 0.0 if sl is None else np.asarray(sl).T.ravel()
        return z

    def unpack(self, z):
        NS, N = self.NS, self.N
        xp = z[self.oX:self.oX + 4 * NS].reshape(NS, 4).T.copy()
        ts = np.ones(NS) if self.fixTime else z[self.oT:self.oT + NS].copy()
        up = z[self.oU:self.oU + 2 * N].reshape(N, 2).T.copy()
        lp = z[self.oL:self.oL + self.V * NS].reshape(NS, self.V).T.copy()
        np_ = z[self.oN:self.oN + 4 * self.nOb * NS].reshape(NS, 4 * self.nOb).T.copy()
        sl = z[self.oS:].reshape(NS, self.nOb).T.copy() if self.variant == "sd" else None
        return xp, up, ts, lp, np_, sl


_TEMPLATES: dict = {}


def _tmpl(key, builder):
    if key not in _TEMPLATES:
        _TEMPLATES[key] = builder()
    return _TEMPLATES[key]


def _dyn_templates(fix):
    """ParkingSignedDist.jl:142-150 (x[.,i+1] == x[.,i] + ...), residual = lhs - rhs."""
    X, Y, psi, v, de, a, ts, Xn, Yn, psin, vn, Ts, L = sp.symbols("X Y psi v de a ts Xn Yn psin vn Ts L")
    h = Ts if fix else ts * Ts
    e = [
        Xn - (X + h * (v + h / 2 * a) * sp.cos(psi + h / 2 * v * sp.tan(de) / L)),
        Yn - (Y + h * (v + h / 2 * a) * sp.sin(psi + h / 2 * v * sp.tan(de) / L)),
        psin - (psi + h * (v + h / 2 * a) * sp.tan(de) / L),
        vn - (v + h * a),
    ]
    vs = [X, Y, psi, v, de, a] + ([] if fix else [ts]) + [Xn, Yn, psin, vn]
    return [Template(ei, vs, [Ts, L]) for ei in e]


def _obca_templates(v, variant):
    """ParkingSignedDist.jl:190-207 for an obstacle with v half-spaces."""
    X, Y, psi, sl, off = sp.symbols("X Y psi sl off")
    lam = sp.symbols(f"lam0:{v}")
    mu = sp.symbols("mu0:4")
    a1 = sp.symbols(f"aa0:{v}")
    a2 = sp.symbols(f"ab0:{v}")
    bb = sp.symbols(f"bb0:{v}")
    g = sp.symbols("g0:4")
    p1 = sum(a1[k] * lam[k] for k in range(v))
    p2 = sum(a2[k] * lam[k] for k in range(v))
    norm = p1 ** 2 + p2 ** 2                                      # :198 (== 1) / ParkingDist.jl:200 (<= 1)
    rot1 = (mu[0] - mu[2]) + sp.cos(psi) * p1 + sp.sin(psi) * p2  # :201
    rot2 = (mu[1] - mu[3]) - sp.sin(psi) * p1 + sp.cos(psi) * p2  # :202
    dist = (-sum(g[k] * mu[k] for k in range(4)) + (X + sp.cos(psi) * off) * p1
            + (Y + sp.sin(psi) * off) * p2 - sum(bb[k] * lam[k] for k in range(v)))  # :205-206
    vs = [X, Y, psi] + list(lam) + list(mu)
    par = list(a1) + list(a2) + list(bb) + list(g) + [off]
    if variant == "sd":
        tn = Template(norm - 1, vs, par)
        td = Template(dist + sl, vs + [sl], par)
    else:
        tn = Template(norm, vs, par)
        td = Template(dist, vs, par)
    return tn, Template(rot1, vs, par), Template(rot2, vs, par), td


def build_parking_nlp(x0, xF, N, Ts, L, ego, XYbounds, nOb, vOb, A, b, rx, ry, ryaw, fixTime,
                      variant="sd"):
    """Same 15 leading arguments as ParkingSignedDist.jl:29 (xWS/uWS are initial values, not model data)."""
    x0 = np.asarray(x0, float).ravel(); xF = np.asarray(xF, float).ravel()
    ego = np.asarray(ego, float).ravel(); XYb = np.asarray(XYbounds, float).ravel()
    A = np.asarray(A, float).reshape(-1, 2); b = np.asarray(b, float).ravel()
    rx = np.asarray(rx, float).ravel(); ry = np.asarray(ry, float).ravel(); ryaw = np.asarray(ryaw, float).ravel()
    lay = Layout(N, nOb, vOb, fixTime, variant)
    NS = lay.NS
    nlp = SparseNLP(lay.n)
    nlp.lay = lay
    fix = bool(fixTime)
    ks = np.arange(NS); kN = np.arange(N)

    # ---------------- objective (ParkingSignedDist.jl:77-93 / ParkingDist.jl:78-94) ----------------
    wa = 0.5 if (fix or variant == "d") else 0.1           # SD var-time 0.1 (:86); SD fix 0.5 (:79); Dist 0.5 (:79,:87)
    wyaw = 0.01 if fix else 0.0001                          # :82 vs :91
    de, a, de2, a2, ts, v, X, Y, psi, sl, r1, r2, r3, TsS = sp.symbols("de a de2 a2 ts v X Y psi sl r1 r2 r3 Ts")
    t_u = _tmpl(("ucost", wa), lambda: Template(0.01 * de ** 2 + wa * a ** 2, [de, a], []))
    nlp.add_obj(Family("u_cost", t_u, np.stack([lay.u(0, kN), lay.u(1, kN)], 1)))
    if fix:
        t_du = _tmpl("du_fix", lambda: Template(0.1 * ((de2 - de) / TsS) ** 2 + 0.1 * ((a2 - a) / TsS) ** 2,
                                                [de, a, de2, a2], [TsS]))
        k = np.arange(N - 1)
        nlp.add_obj(Family("du_cost", t_du, np.stack([lay.u(0, k), lay.u(1, k), lay.u(0, k + 1), lay.u(1, k + 1)], 1),
                           np.full((N - 1, 1), Ts)))
        t_du0 = _tmpl("du0_fix", lambda: Template(0.1 * (de / TsS) ** 2 + 0.1 * (a / TsS) ** 2, [de, a], [TsS]))
        nlp.add_obj(Family("du0_cost", t_du0, [[lay.u(0, 0), lay.u(1, 0)]], [[Ts]]))
    else:
        t_du = _tmpl("du_var", lambda: Template(0.1 * ((de2 - de) / (ts * TsS)) ** 2 + 0.1 * ((a2 - a) / (ts * TsS)) ** 2,
                                                [de, a, de2, a2, ts], [TsS]))
        k = np.arange(N - 1)
        nlp.add_obj(Family("du_cost", t_du,
                           np.stack([lay.u(0, k), lay.u(1, k), lay.u(0, k + 1), lay.u(1, k + 1), lay.ts(k)], 1),
                           np.full((N - 1, 1), Ts)))
        t_du0 = _tmpl("du0_var", lambda: Template(0.1 * (de / (ts * TsS)) ** 2 + 0.1 * (a / (ts * TsS)) ** 2,
                                                  [de, a, ts], [TsS]))
        nlp.add_obj(Family("du0_cost", t_du0, [[lay.u(0, 0), lay.u(1, 0), lay.ts(0)]], [[Ts]]))
        t_ts = _tmpl("ts_cost", lambda: Template(0.5 * ts + 1 * ts ** 2, [ts], []))
        nlp.add_obj(Family("ts_cost", t_ts, lay.ts(ks).reshape(-1, 1)))
    t_v = _tmpl("v_cost", lambda: Template(0.0001 * v ** 2, [v], []))
    nlp.add_obj(Family("v_cost", t_v, lay.x(3, ks).reshape(-1, 1)))
    t_tr = _tmpl(("track", wyaw), lambda: Template(0.001 * (X - r1) ** 2 + 0.001 * (Y - r2) ** 2 + wyaw * (psi - r3) ** 2,
                                                   [X, Y, psi], [r1, r2, r3]))
    nlp.add_obj(Family("track", t_tr, np.stack([lay.x(0, ks), lay.x(1, ks), lay.x(2, ks)], 1),
                       np.stack([rx, ry, ryaw], 1)))
    if variant == "sd":
        t_sl = _tmpl("sl_cost", lambda: Template(1e2 * sl + 1e4 * sl ** 2, [sl], []))
        jj, kk = np.meshgrid(np.arange(nOb), ks, indexing="ij")
        nlp.add_obj(Family("sl_cost", t_sl, lay.sl(jj.ravel(), kk.ravel()).reshape(-1, 1)))

    # ---------------- simple bounds (written as @constraint rows in the reference, :100-115) ----------------
    nlp.zL[lay.u(0, kN)] = -0.6; nlp.zU[lay.u(0, kN)] = 0.6
    nlp.zL[lay.u(1, kN)] = -0.4; nlp.zU[lay.u(1, kN)] = 0.4
    nlp.zL[lay.x(0, ks)] = XYb[0]; nlp.zU[lay.x(0, ks)] = XYb[1]
    nlp.zL[lay.x(1, ks)] = XYb[2]; nlp.zU[lay.x(1, ks)] = XYb[3]
    nlp.zL[lay.x(3, ks)] = -1.0; nlp.zU[lay.x(3, ks)] = 2.0
    if not fix:
        nlp.zL[lay.ts(ks)] = 0.8; nlp.zU[lay.ts(ks)] = 1.2
    nlp.zL[lay.oL:lay.oN] = 0.0          # l .>= 0
    nlp.zL[lay.oN:lay.oS] = 0.0          # n .>= 0   (sl is free)

    # ---------------- start / end (:122-131) ----------------
    q, c0 = sp.symbols("q c0")
    t_fixv = _tmpl("fixv", lambda: Template(q - c0, [q], [c0]))
    nlp.add_eq(Family("start", t_fixv, lay.x(np.arange(4), 0).reshape(-1, 1), x0.reshape(-1, 1)))
    nlp.add_eq(Family("end", t_fixv, lay.x(np.arange(4), N).reshape(-1, 1), xF.reshape(-1, 1)))

    # ---------------- dynamics + timeScale chain (:139-155) ----------------
    tdyn = _tmpl(("dyn", fix), lambda: _dyn_templates(fix))
    cols = [lay.x(0, kN), lay.x(1, kN), lay.x(2, kN), lay.x(3, kN), lay.u(0, kN), lay.u(1, kN)]
    if not fix:
        cols.append(lay.ts(kN))
    cols += [lay.x(0, kN + 1), lay.x(1, kN + 1), lay.x(2, kN + 1), lay.x(3, kN + 1)]
    idx = np.stack(cols, 1)
    par = np.tile([Ts, L], (N, 1))
    for i in range(4):
        nlp.add_eq(Family(f"dyn{i}", tdyn[i], idx, par))
    if not fix:
        t1, t2 = sp.symbols("t1 t2")
        t_ch = _tmpl("chain", lambda: Template(t1 - t2, [t1, t2], []))
        nlp.add_eq(Family("chain", t_ch, np.stack([lay.ts(kN), lay.ts(kN + 1)], 1)))

    # ---------------- steering-rate rows (:157-174) ----------------
    dp = sp.symbols("dp")
    k1 = np.arange(1, N)
    if fix:
        t_r0 = _tmpl("rate0_fix", lambda: Template((0 - de) / TsS, [de], [TsS]))
        t_r = _tmpl("rate_fix", lambda: Template((dp - de) / TsS, [dp, de], [TsS]))
        nlp.add_ineq(Family("rate0", t_r0, [[lay.u(0, 0)]], [[Ts]], lo=-0.6, hi=0.6))
        nlp.add_ineq(Family("rate", t_r, np.stack([lay.u(0, k1 - 1), lay.u(0, k1)], 1), np.full((N - 1, 1), Ts),
                            lo=-0.6, hi=0.6))
    else:
        t_r0 = _tmpl("rate0_var", lambda: Template((0 - de) / (ts * TsS), [de, ts], [TsS]))
        t_r = _tmpl("rate_var", lambda: Template((dp - de) / (ts * TsS), [dp, de, ts], [TsS]))
        nlp.add_ineq(Family("rate0", t_r0, [[lay.u(0, 0), lay.ts(0)]], [[Ts]], lo=-0.6, hi=0.6))
        nlp.add_ineq(Family("rate", t_r, np.stack([lay.u(0, k1 - 1), lay.u(0, k1), lay.ts(k1)], 1),
                            np.full((N - 1, 1), Ts), lo=-0.6, hi=0.6))

    # ---------------- OBCA rows (:182-208) ----------------
    W_ev = ego[1] + ego[3]; L_ev = ego[0] + ego[2]
    g = np.array([L_ev / 2, W_ev / 2, L_ev / 2, W_ev / 2])
    offset = (ego[0] + ego[2]) / 2 - ego[2]
    for j in range(nOb):
        vj = lay.vOb[j]
        tn, tr1, tr2, td = _tmpl(("obca", vj, variant), lambda: _obca_templates(vj, variant))
        r0 = lay.voff[j]
        cols = [lay.x(0, ks), lay.x(1, ks), lay.x(2, ks)] + [lay.l(r0 + r, ks) for r in range(vj)] \
            + [lay.mu(4 * j + r, ks) for r in range(4)]
        idx = np.stack(cols, 1)
        par = np.tile(np.concatenate([A[r0:r0 + vj, 0], A[r0:r0 + vj, 1], b[r0:r0 + vj], g, [offset]]), (NS, 1))
        if variant == "sd":
            nlp.add_eq(Family(f"norm{j}", tn, idx, par))
        else:
            nlp.add_ineq(Family(f"norm{j}", tn, idx, par, hi=1.0))
        nlp.add_eq(Family(f"rot1_{j}", tr1, idx, par))
        nlp.add_eq(Family(f"rot2_{j}", tr2, idx, par))
        if variant == "sd":
            idxd = np.concatenate([idx, lay.sl(j, ks).reshape(-1, 1)], 1)
        else:
            idxd = idx
        nlp.add_ineq(Family(f"dist{j}", td, idxd, par, lo=DMIN))
    return nlp


def initial_point(lay: Layout, xWS, uWS, lWS, nWS):
    """ParkingSignedDist.jl:213-222: ts=1, x=xWS', u=uWS[1:N,:]', l=lWS', n=nWS', sl=0 (JuMP default)."""
    xWS = np.asarray(xWS, float); uWS = np.asarray(uWS, float)
    z = np.zeros(lay.n)
    z[lay.oX:lay.oX + 4 * lay.NS] = xWS[:lay.NS, :4].ravel()
    if not lay.fixTime:
        z[lay.oT:lay.oT + lay.NS] = 1.0
    z[lay.oU:lay.oU + 2 * lay.N] = uWS[:lay.N, :2].ravel()
    z[lay.oL:lay.oL + lay.V * lay.NS] = np.asarray(lWS, float)[:lay.NS, :lay.V].ravel()
    z[lay.oN:lay.oN + 4 * lay.nOb * lay.NS] = np.asarray(nWS, float)[:lay.NS, :4 * lay.nOb].ravel()
    return z
