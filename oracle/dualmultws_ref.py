"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of AutonomousParking/DualMultWS.jl:29-86.  The reference model is
separable: no row couples two different (stage i, obstacle j) pairs
(DualMultWS.jl:57-75), so it is solved here as (N+1)*nOb independent programs

    max  d = -g'mu + (A_j t_i - b_j)'lam                      (:72-73, objective :52)
    s.t. (A_j'lam)_1^2 + (A_j'lam)_2^2 <= 1                   (:65)
         mu1-mu3 + c*p1 + s*p2 = 0, mu2-mu4 - s*p1 + c*p2 = 0  (:68-69)
         lam >= 0, mu >= 0                                     (:54-55)

two ways: (a) scipy SLSQP on the restated program; (b) closed-form geometry --
the optimum d* equals the Euclidean distance between the ego rectangle at pose
i and polyhedron j (0 when they overlap), SURVEY.md section 8c.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize


def ego_geometry(ego):
    ego = np.asarray(ego, float).ravel()
    W_ev = ego[1] + ego[3]; L_ev = ego[0] + ego[2]                 # DualMultWS.jl:39-40
    g = np.array([L_ev / 2, W_ev / 2, L_ev / 2, W_ev / 2])         # :42
    offset = (ego[0] + ego[2]) / 2 - ego[2]                         # :45
    return g, offset


def solve_one(Aj, bj, pose, g, offset):
    """SLSQP solve of one (i,j) program; returns lam, mu, d."""
    v = Aj.shape[0]
    X, Y, psi = pose
    c, s = np.cos(psi), np.sin(psi)
    tc = np.array([X + c * offset, Y + s * offset])
    rho = Aj @ tc - bj

    def unpack(q):
        return q[:v], q[v:v + 4]

    def negd(q):
        lam, mu = unpack(q)
        return -(-g @ mu + rho @ lam)

    def negd_grad(q):
        return -np.concatenate([rho, -g])

    def eq(q):
        lam, mu = unpack(q)
        p = Aj.T @ lam
        return np.array([mu[0] - mu[2] + c * p[0] + s * p[1], mu[1] - mu[3] - s * p[0] + c * p[1]])

    def ineq(q):
        lam, _ = unpack(q)
        p = Aj.T @ lam
        return 1.0 - p @ p

    best = None
    for q0 in (np.zeros(v + 4), np.full(v + 4, 0.3)):
        r = minimize(negd, q0, jac=negd_grad, method="SLSQP", bounds=[(0, None)] * (v + 4),
                     constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                     options=dict(ftol=1e-14, maxiter=500))
        if best is None or r.fun < best.fun - 1e-12:
            best = r
    lam, mu = unpack(best.x)
    return lam, mu, -best.fun


def dualmultws(N, nOb, vOb, A, b, rx, ry, ryaw, ego):
    """Same call surface as DualMultWS.jl:29 (+ explicit ego, a global in the reference, :39).
    Returns lp (N+1)xsum(vOb), np (N+1)x4nOb (already transposed like :81-84) and d (N+1)xnOb."""
    vOb = [int(x) for x in np.asarray(vOb).ravel()]
    A = np.asarray(A, float).reshape(-1, 2); b = np.asarray(b, float).ravel()
    g, offset = ego_geometry(ego)
    V = sum(vOb); off = np.concatenate([[0], np.cumsum(vOb)]).astype(int)
    lp = np.zeros((N + 1, V)); npp = np.zeros((N + 1, 4 * nOb)); d = np.zeros((N + 1, nOb))
    for i in range(N + 1):
        for j in range(nOb):
            lam, mu, dd = solve_one(A[off[j]:off[j + 1]], b[off[j]:off[j + 1]], (rx[i], ry[i], ryaw[i]), g, offset)
            lp[i, off[j]:off[j + 1]] = lam; npp[i, 4 * j:4 * j + 4] = mu; d[i, j] = dd
    return lp, npp, d


# ------------------------------------------------------------------------------------------
# closed-form check: distance between the ego rectangle and {y : Aj y <= bj}
# ------------------------------------------------------------------------------------------
def _rect_corners(pose, g, offset):
    X, Y, psi = pose
    c, s = np.cos(psi), np.sin(psi)
    ctr = np.array([X + c * offset, Y + s * offset])
    R = np.array([[c, -s], [s, c]])
    loc = np.array([[g[0], g[1]], [-g[2], g[1]], [-g[2], -g[3]], [g[0], -g[3]]])
    return ctr + loc @ R.T


def _pt_seg(p, a, b_):
    ab = b_ - a
    t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-300), 0.0, 1.0)
    return np.linalg.norm(p - (a + t * ab))


def _pt_poly_dist(p, Aj, bj, big=1e4):
    """distance from point to polyhedron {A y <= b} with <=2 rows (half-plane or wedge) or general via projection."""
    viol = Aj @ p - bj
    if (viol <= 0).all():
        return 0.0
    v = Aj.shape[0]
    cands = []
    for r in range(v):                                   # project on each face, keep if feasible
        a = Aj[r]; q = p - a * (a @ p - bj[r]) / (a @ a)
        if (Aj @ q - bj <= 1e-9).all():
            cands.append(np.linalg.norm(p - q))
    for r1 in range(v):                                  # vertices
        for r2 in range(r1 + 1, v):
            M = Aj[[r1, r2]]
            if abs(np.linalg.det(M)) > 1e-12:
                q = np.linalg.solve(M, bj[[r1, r2]])
                if (Aj @ q - bj <= 1e-9).all():
                    cands.append(np.linalg.norm(p - q))
    return min(cands)


def rect_poly_distance(pose, Aj, bj, g, offset):
    """min distance rectangle <-> polyhedron; 0 when intersecting.  Vertex-edge enumeration."""
    C = _rect_corners(pose, g, offset)
    # intersection test by sampling the rectangle boundary + vertices of the polyhedron inside the rectangle
    best = min(_pt_poly_dist(c_, Aj, bj) for c_ in C)
    # polyhedron vertices against rectangle edges
    v = Aj.shape[0]
    verts = []
    for r1 in range(v):
        for r2 in range(r1 + 1, v):
            M = Aj[[r1, r2]]
            if abs(np.linalg.det(M)) > 1e-12:
                q = np.linalg.solve(M, bj[[r1, r2]])
                if (Aj @ q - bj <= 1e-9).all():
                    verts.append(q)
    X, Y, psi = pose
    c, s = np.cos(psi), np.sin(psi)
    ctr = np.array([X + c * offset, Y + s * offset])
    R = np.array([[c, -s], [s, c]])
    for q in verts:
        ql = R.T @ (q - ctr)
        inside = (-g[2] <= ql[0] <= g[0]) and (-g[3] <= ql[1] <= g[1])
        if inside:
            return 0.0
        for e in range(4):
            best = min(best, _pt_seg(q, C[e], C[(e + 1) % 4]))
    # edges crossing: sample rectangle edges densely for feasibility (cheap and adequate for <=2 rows)
    for e in range(4):
        for t in np.linspace(0, 1, 33):
            p = C[e] + t * (C[(e + 1) % 4] - C[e])
            if (Aj @ p - bj <= 0).all():
                return 0.0
    return best


# ------------------------------------------------------------------------------------------
# (c) the reference's way: ONE model with all (i, j) blocks, solved by the IPOPT stand-in
# ------------------------------------------------------------------------------------------
_DW_T = {}


def _dw_templates(v=None):
    """sympy templates of DualMultWS.jl:52-73 (objective term; norm / rot1 / rot2 / d-definition rows of a block with v half-spaces)."""
    import sympy as sp
    from .sparse_nlp import Template
    dsym = sp.Symbol("d")
    if "obj" not in _DW_T:
        _DW_T["obj"] = Template(-dsym, [dsym], [])
    if v is not None and ("blk", v) not in _DW_T:
        lam = sp.symbols(f"lam0:{v}"); mu = sp.symbols("mu0:4")
        a1 = sp.symbols(f"aa0:{v}"); a2 = sp.symbols(f"ab0:{v}"); rho = sp.symbols(f"rho0:{v}")
        gs = sp.symbols("g0:4"); cs, ss = sp.symbols("cs ss")
        p1 = sum(a1[i] * lam[i] for i in range(v)); p2 = sum(a2[i] * lam[i] for i in range(v))
        vs = list(lam) + list(mu); par = list(a1) + list(a2) + list(rho) + list(gs) + [cs, ss]
        _DW_T[("blk", v)] = (Template(p1 ** 2 + p2 ** 2, vs, par),
                             Template((mu[0] - mu[2]) + cs * p1 + ss * p2, vs, par),
                             Template((mu[1] - mu[3]) - ss * p1 + cs * p2, vs, par),
                             Template(dsym - (-sum(gs[i] * mu[i] for i in range(4)) + sum(rho[i] * lam[i] for i in range(v))),
                                      vs + [dsym], par))
    return _DW_T


def build_dualmultws_nlp(N, nOb, vOb, A, b, rx, ry, ryaw, ego):
    """DualMultWS.jl:36-77 as one sparse NLP (variables l, n, d; Max sum(d)).  Returns (nlp, banded symmetric ordering, (oN, oD))."""
    from .sparse_nlp import Family, SparseNLP
    vOb = [int(x) for x in np.asarray(vOb).ravel()]
    A = np.asarray(A, float).reshape(-1, 2); b = np.asarray(b, float).ravel()
    g, offset = ego_geometry(ego)
    NS = N + 1; V = sum(vOb); voff = np.concatenate([[0], np.cumsum(vOb)]).astype(int)
    oN = V * NS; oD = oN + 4 * nOb * NS
    n = oD + nOb * NS
    nlp = SparseNLP(n)
    ks = np.arange(NS)
    c, s = np.cos(ryaw), np.sin(ryaw)
    tcx, tcy = np.asarray(rx) + c * offset, np.asarray(ry) + s * offset
    _dw_templates()
    jj, kk = np.meshgrid(np.arange(nOb), ks, indexing="ij")
    nlp.add_obj(Family("negd", _DW_T["obj"], (oD + nOb * kk.ravel() + jj.ravel()).reshape(-1, 1)))
    nlp.zL[:oD] = 0.0
    for j in range(nOb):
        v = vOb[j]
        _dw_templates(v)
        tn, t1, t2, td = _DW_T[("blk", v)]
        r0 = voff[j]
        idx = np.stack([V * ks + r0 + r for r in range(v)] + [oN + 4 * nOb * ks + 4 * j + m for m in range(4)], 1)
        rho_ = np.stack([A[r0 + r, 0] * tcx + A[r0 + r, 1] * tcy - b[r0 + r] for r in range(v)], 1)
        par = np.concatenate([np.tile(A[r0:r0 + v, 0], (NS, 1)), np.tile(A[r0:r0 + v, 1], (NS, 1)), rho_,
                              np.tile(g, (NS, 1)), c[:, None], s[:, None]], 1)
        nlp.add_ineq(Family(f"norm{j}", tn, idx, par, hi=1.0))
        nlp.add_eq(Family(f"rot1_{j}", t1, idx, par))
        nlp.add_eq(Family(f"rot2_{j}", t2, idx, par))
        nlp.add_eq(Family(f"ddef{j}", td, np.concatenate([idx, (oD + nOb * ks + j).reshape(-1, 1)], 1), par))
    # banded symmetric ordering: (l_k, n_k) -> equality rows of stage k -> d_k  (d has no diagonal of its own, so
    # it must follow the row that defines it for a pivot-free LDL')
    vkey = np.zeros(n)
    vkey[:oN] = 3 * np.repeat(ks, V); vkey[oN:oD] = 3 * np.repeat(ks, 4 * nOb); vkey[oD:] = 3 * np.repeat(ks, nOb) + 2
    rkey = np.zeros(nlp.mE)
    for fam in nlp.eq:
        rkey[fam.row0:fam.row0 + fam.n] = 3 * ks + 1
    order = np.argsort(np.concatenate([vkey, rkey]), kind="stable")
    return nlp, order, (oN, oD)


def dualmultws_ipm(N, nOb, vOb, A, b, rx, ry, ryaw, ego, opts=None):
    """build_dualmultws_nlp solved by oracle/ipm_ref.py with the reference's options (tol=1e-5, max_iter=100, DualMultWS.jl:36-37).
    Returns lp (N+1)xV, np (N+1)x4nOb, d (N+1)xnOb."""
    from . import ipm_ref
    nlp, order, (oN, oD) = build_dualmultws_nlp(N, nOb, vOb, A, b, rx, ry, ryaw, ego)
    n = nlp.n; NS = N + 1; V = int(np.sum(np.asarray(vOb)))
    o = opts or ipm_ref.IpmOptions(tol=1e-5, max_iter=100, linsolve="sparse")
    if o.linsolve == "sparse" and o.order is None:
        o.order = order
    res = ipm_ref.solve(nlp, np.zeros(n), o)          # JuMP default start = 0 (DualMultWS.jl has no setvalue)
    z = res.z
    return z[:oN].reshape(NS, V), z[oN:oD].reshape(NS, 4 * nOb), z[oD:].reshape(NS, nOb), res
