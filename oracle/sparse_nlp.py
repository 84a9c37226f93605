"""ORACLE (test infrastructure only -- never imported by the product path).

Generic sparse NLP container whose derivatives come from sympy, NOT from the
hand-derived formulas used in the CUDA kernels.  A problem is a list of
"families": one symbolic template expression instantiated many times (once per
stage / per obstacle) on different global variable indices and parameters.
This mirrors how JuMP stores the reference model (one expression tree per
@NLconstraint instance, e.g. ParkingSignedDist.jl:190-208) and gives exact
first/second derivatives that are independent of the product code.

    min f(z)   s.t.  cE(z) = 0,   gL <= g(z) <= gU,   zL <= z <= zU
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import sympy as sp


class Template:
    """value / gradient / lower-triangular Hessian of one symbolic expression."""

    _cache: dict = {}

    def __init__(self, expr, vars_, params):
        self.nv = len(vars_)
        self.np_ = len(params)
        grad = [sp.diff(expr, v) for v in vars_]
        self.hpairs = []
        hexpr = []
        for i in range(self.nv):
            for j in range(i + 1):
                h = sp.diff(grad[i], vars_[j])
                if h != 0:
                    self.hpairs.append((i, j))
                    hexpr.append(h)
        args = list(vars_) + list(params)
        # kept for the C code generator of the compiled CPU baseline (oracle/cpu_ipm/gen.py)
        self.expr, self.vars_, self.params, self.grad_expr, self.hess_expr = expr, list(vars_), list(params), grad, hexpr
        self._f = sp.lambdify(args, expr, "numpy", cse=True)
        self._g = sp.lambdify(args, grad, "numpy", cse=True)
        self._h = sp.lambdify(args, hexpr, "numpy", cse=True) if hexpr else None
        self.gnz = [i for i in range(self.nv) if grad[i] != 0]

    @staticmethod
    def _stack(vals, n):
        return np.stack([np.broadcast_to(np.asarray(v, dtype=float), (n,)) for v in vals], axis=1) \
            if len(vals) else np.zeros((n, 0))

    def f(self, V, P):
        n = V.shape[0]
        return np.broadcast_to(np.asarray(self._f(*V.T, *P.T), dtype=float), (n,)).copy()

    def g(self, V, P):
        return self._stack(self._g(*V.T, *P.T), V.shape[0])

    def h(self, V, P):
        if self._h is None:
            return np.zeros((V.shape[0], 0))
        return self._stack(self._h(*V.T, *P.T), V.shape[0])


class Family:
    def __init__(self, name, tmpl, idx, params=None, lo=None, hi=None):
        self.name = name
        self.t = tmpl
        self.idx = np.asarray(idx, dtype=np.int64).reshape(-1, tmpl.nv)
        self.n = self.idx.shape[0]
        self.P = np.zeros((self.n, 0)) if params is None else np.asarray(params, dtype=float).reshape(self.n, -1)
        assert self.P.shape[1] == tmpl.np_, (name, self.P.shape, tmpl.np_)
        self.lo = None if lo is None else np.broadcast_to(np.asarray(lo, dtype=float), (self.n,)).copy()
        self.hi = None if hi is None else np.broadcast_to(np.asarray(hi, dtype=float), (self.n,)).copy()
        self.row0 = 0

    def val(self, z):
        return self.t.f(z[self.idx], self.P)

    def grad(self, z):
        return self.t.g(z[self.idx], self.P)

    def hess(self, z):
        return self.t.h(z[self.idx], self.P)


class SparseNLP:
    def __init__(self, n, names=None):
        self.n = n
        self.obj: list[Family] = []
        self.eq: list[Family] = []
        self.ineq: list[Family] = []
        self.zL = np.full(n, -np.inf)
        self.zU = np.full(n, np.inf)
        self.names = names

    # -- construction ------------------------------------------------------
    def add_obj(self, fam):
        self.obj.append(fam)

    def add_eq(self, fam):
        fam.row0 = self.mE
        self.eq.append(fam)

    def add_ineq(self, fam):
        fam.row0 = self.mI
        self.ineq.append(fam)

    @property
    def mE(self):
        return sum(f.n for f in self.eq)

    @property
    def mI(self):
        return sum(f.n for f in self.ineq)

    @property
    def gL(self):
        return np.concatenate([f.lo if f.lo is not None else np.full(f.n, -np.inf) for f in self.ineq]) \
            if self.ineq else np.zeros(0)

    @property
    def gU(self):
        return np.concatenate([f.hi if f.hi is not None else np.full(f.n, np.inf) for f in self.ineq]) \
            if self.ineq else np.zeros(0)

    # -- evaluation ----------------------------------------------------------
    def f(self, z):
        return float(sum(fam.val(z).sum() for fam in self.obj))

    def grad(self, z):
        g = np.zeros(self.n)
        for fam in self.obj:
            np.add.at(g, fam.idx.ravel(), fam.grad(z).ravel())
        return g

    def _c(self, fams, z):
        return np.concatenate([fam.val(z) for fam in fams]) if fams else np.zeros(0)

    def cE(self, z):
        return self._c(self.eq, z)

    def g(self, z):
        return self._c(self.ineq, z)

    def _jac(self, fams, m, z):
        rows, cols, vals = [], [], []
        for fam in fams:
            G = fam.grad(z)
            nz = fam.t.gnz
            r = np.repeat(fam.row0 + np.arange(fam.n), len(nz))
            rows.append(r)
            cols.append(fam.idx[:, nz].ravel())
            vals.append(G[:, nz].ravel())
        if not rows:
            return sps.csr_matrix((m, self.n))
        return sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                              shape=(m, self.n)).tocsr()

    def JE(self, z):
        return self._jac(self.eq, self.mE, z)

    def JI(self, z):
        return self._jac(self.ineq, self.mI, z)

    def hess(self, z, yE, yI, sigma=1.0):
        """sigma*Hess f + sum yE_i Hess cE_i + sum yI_i Hess g_i (full symmetric, csr)."""
        rows, cols, vals = [], [], []

        def acc(fam, w):
            if not fam.t.hpairs:
                return
            Hv = fam.hess(z) * w[:, None]
            for q, (i, j) in enumerate(fam.t.hpairs):
                rows.append(fam.idx[:, i]); cols.append(fam.idx[:, j]); vals.append(Hv[:, q])

        for fam in self.obj:
            acc(fam, np.full(fam.n, sigma))
        for fam in self.eq:
            acc(fam, yE[fam.row0:fam.row0 + fam.n])
        for fam in self.ineq:
            acc(fam, yI[fam.row0:fam.row0 + fam.n])
        if not rows:
            return sps.csr_matrix((self.n, self.n))
        r = np.concatenate(rows); c = np.concatenate(cols); v = np.concatenate(vals)
        lo = np.minimum(r, c); hi = np.maximum(r, c)            # store lower triangle: (hi, lo)
        L = sps.coo_matrix((v, (hi, lo)), shape=(self.n, self.n)).tocsr()
        D = sps.diags(L.diagonal())
        return (L + L.T - D).tocsr()

    # -- structure statistics (SURVEY A.6 cross-check) ------------------------
    def nnz_jac(self):
        return sum(f.n * len(f.t.gnz) for f in self.eq + self.ineq)
