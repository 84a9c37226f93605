"""ORACLE / CPU BASELINE (test infrastructure only).  ctypes front end of the compiled generic interior-point solver
(cpu_ipm.cpp: the oracle's IPOPT stand-in oracle/ipm_ref.py restated in C++, derivatives generated from the oracle's sympy
templates by gen.py, skyline LDL' of the full augmented system).  Used by tests/ and by bench.py's cpu_baseline /
--impl reference legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcpu_ipm.so")


class Opts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double), ("mu_min_factor", C.c_double),
                ("kappa_eps", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double), ("tau_min", C.c_double),
                ("kappa1", C.c_double), ("kappa2", C.c_double), ("kappa_sigma", C.c_double), ("s_max", C.c_double),
                ("dual_inf_tol", C.c_double), ("constr_viol_tol", C.c_double), ("compl_inf_tol", C.c_double),
                ("dw_min", C.c_double), ("dw_first", C.c_double), ("dw_max", C.c_double), ("kw_minus", C.c_double),
                ("kw_plus", C.c_double), ("kw_plus_first", C.c_double), ("gamma_theta", C.c_double), ("gamma_phi", C.c_double),
                ("delta", C.c_double), ("s_theta", C.c_double), ("s_phi", C.c_double), ("eta_phi", C.c_double),
                ("gamma_alpha", C.c_double), ("max_backtrack", C.c_int), ("dc_value", C.c_double), ("max_kick", C.c_int),
                ("freeze_degenerate", C.c_double)]


def build(force=False):
    """Generate templates_gen.h from the oracle's sympy templates and compile libcpu_ipm.so (g++ -O2 -fopenmp)."""
    from . import gen
    srcs = [os.path.join(HERE, "cpu_ipm.cpp"), os.path.join(HERE, "gen.py"), os.path.join(HERE, "..", "parking_nlp.py"),
            os.path.join(HERE, "..", "sparse_nlp.py"), os.path.join(HERE, "..", "dualmultws_ref.py"),
            os.path.join(HERE, "..", "quadcopter_nlp.py")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        gen.generate()
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", SO, os.path.join(HERE, "cpu_ipm.cpp")])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.cpu_ipm_template_name.restype = C.c_char_p
        assert _lib.cpu_ipm_opts_size() == C.sizeof(Opts)
    return _lib


def default_opts(o=None):
    """The options of oracle/ipm_ref.IpmOptions (= the Ipopt options of ParkingSignedDist.jl:41-43 + Ipopt defaults)."""
    from .. import ipm_ref
    o = o or ipm_ref.IpmOptions()
    c = Opts()
    for name, _ in Opts._fields_:
        setattr(c, name, getattr(o, name))
    return c


def export(nlp):
    """Structure of a SparseNLP of oracle/parking_nlp.py for the C solver + this problem's parameter blob."""
    from . import gen
    ids, _ = gen.template_ids()
    names = {i: lib().cpu_ipm_template_name(i).decode() for i in range(lib().cpu_ipm_n_templates())}
    desc, idx, par, lo, hi = [], [], [], [], []
    io = po = bo = 0
    for kind, fams in ((0, nlp.obj), (1, nlp.eq), (2, nlp.ineq)):
        for f in fams:
            tid, name = ids[id(f.t)]
            assert names[tid] == name, (names[tid], name)        # the compiled header and the live templates agree
            b_off = -1
            if kind == 2:
                b_off = bo
                lo.append(np.where(np.isfinite(f.lo), f.lo, -1e300) if f.lo is not None else np.full(f.n, -1e300))
                hi.append(np.where(np.isfinite(f.hi), f.hi, 1e300) if f.hi is not None else np.full(f.n, 1e300))
                bo += f.n
            desc.append([kind, tid, f.n, io, po, b_off])
            idx.append(f.idx.astype(np.int32).ravel()); par.append(np.asarray(f.P, float).ravel())
            io += f.idx.size; po += f.P.size
    st = dict(n=nlp.n, mE=nlp.mE, mI=nlp.mI, desc=np.asarray(desc, np.int32), idx=np.concatenate(idx),
              lo=np.concatenate(lo) if lo else np.zeros(1), hi=np.concatenate(hi) if hi else np.zeros(1),
              zL=np.where(np.isfinite(nlp.zL), nlp.zL, -1e300), zU=np.where(np.isfinite(nlp.zU), nlp.zU, 1e300))
    return st, (np.concatenate(par) if par else np.zeros(0))


class Prepared:
    """Exported model structure + parameter blobs of a list of problems that share one structure (the model-build part of a call)."""

    def __init__(self, nlps, dc_rows, order):
        self.nlps = nlps
        self.st, _ = export(nlps[0])
        self.pars = np.ascontiguousarray(np.stack([export(m)[1] for m in nlps]), float)
        self.dc = np.ascontiguousarray(dc_rows, np.uint8)
        self.order = np.ascontiguousarray(order, np.int32)


def run(prep: Prepared, z0s, opts=None, nthreads=1):
    """Solve the prepared problems from the starting points z0s.  Returns dict(z (B,n), status, iters, err, seconds (per problem),
    setup_seconds (symbolic skyline structure))."""
    st = prep.st
    B = len(prep.nlps)
    z0 = np.ascontiguousarray(np.stack(z0s), float)
    o = opts or default_opts()
    z = np.zeros((B, st["n"])); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); err = np.zeros(B); sec = np.zeros(B)
    setup = C.c_double(0.0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().cpu_ipm_solve_batch(C.c_int(st["n"]), C.c_int(st["mE"]), C.c_int(st["mI"]), C.c_int(len(st["desc"])), P(st["desc"]), P(st["idx"]),
                                   P(st["lo"]), P(st["hi"]), P(st["zL"]), P(st["zU"]), P(prep.dc), P(prep.order), C.byref(o), C.c_int(B),
                                   C.c_int(prep.pars.shape[1]), P(prep.pars), P(z0), C.c_int(int(nthreads)), P(z), P(status), P(iters), P(err),
                                   P(sec), C.byref(setup))
    assert rc == 0, rc
    return dict(z=z, status=status, iters=iters, err=err, seconds=sec, setup_seconds=setup.value)


def solve_batch(nlps, z0s, dc_rows, order, opts=None, nthreads=1):
    """nlps: SparseNLPs of ONE structure (same scenario / variant / N; parameters differ), z0s: their starting points."""
    return run(Prepared(nlps, dc_rows, order), z0s, opts, nthreads)


class ParkingCall:
    """The 17-argument call of ParkingSignedDist.jl:29 / ParkingDist.jl:29 for problems `idxs` of a scenario batch
    (obca_b200.scenarios), split the way the reference splits it: model build (here: oracle/parking_nlp.py + the DualMultWS
    model, python) | DualMultWS solve (:219) | solve(m) (:240).  prepare once, run() as often as needed."""

    def __init__(self, sc, idxs, variant="sd", fixTime=0):
        import time
        from ..dualmultws_ref import build_dualmultws_nlp
        from ..parking_nlp import build_parking_nlp
        from ..parking_solve import dc_mask, solver_view, stage_order
        t0 = time.time()
        self.sc, self.idxs, self.N = sc, list(idxs), sc["N"]
        N = self.N; Ts = sc["Ts_fix"] if fixTime else sc["Ts"]
        nlps = [solver_view(build_parking_nlp(sc["x0"][i], sc["xF"], N, Ts, sc["L"], sc["ego"], sc["XYbounds"], sc["nOb"], sc["vOb"], sc["A"],
                                              sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], fixTime, variant)) for i in self.idxs]
        self.main = Prepared(nlps, dc_mask(nlps[0]), stage_order(nlps[0]))
        dws = [build_dualmultws_nlp(N, sc["nOb"], sc["vOb"], sc["A"], sc["b"], sc["rx"][i], sc["ry"][i], sc["ryaw"][i], sc["ego"]) for i in self.idxs]
        self.dw = Prepared([d[0] for d in dws], np.zeros(dws[0][0].mE, bool), dws[0][1])
        self.oN, self.oD = dws[0][2]
        self.wall_build = time.time() - t0

    def run(self, opts=None, nthreads=1, lWS=None, nWS=None):
        import time
        from .. import ipm_ref
        from ..parking_nlp import initial_point
        sc, N = self.sc, self.N
        t0 = time.time()
        r_ws = None
        if lWS is None:
            # DualMultWS.jl:36-77: the reference's second JuMP model (tol 1e-5, max_iter 100, start = 0), same compiled solver
            r_ws = run(self.dw, [np.zeros(self.dw.st["n"])] * len(self.idxs), default_opts(ipm_ref.IpmOptions(tol=1e-5, max_iter=100)), nthreads)
            V = int(np.sum(sc["vOb"]))
            lWS = [r_ws["z"][q][:self.oN].reshape(N + 1, V) for q in range(len(self.idxs))]
            nWS = [r_ws["z"][q][self.oN:self.oD].reshape(N + 1, 4 * sc["nOb"]) for q in range(len(self.idxs))]
        z0s = [initial_point(self.main.nlps[q].lay, sc["xWS"][i], sc["uWS"][i], lWS[q], nWS[q]) for q, i in enumerate(self.idxs)]
        t_ws = time.time() - t0
        t0 = time.time()
        r = run(self.main, z0s, opts, nthreads)
        r["wall_solve"] = time.time() - t0; r["wall_build"] = self.wall_build; r["wall_dualws"] = t_ws
        r["dualws"] = r_ws
        r["out"] = [self.main.nlps[q].lay.unpack(r["z"][q]) for q in range(len(self.idxs))]
        r["nlp0"] = self.main.nlps[0]
        return r


def solve_parking_batch(sc, idxs, variant="sd", fixTime=0, opts=None, nthreads=1, lWS=None, nWS=None):
    """ParkingCall(...).run(...) in one go."""
    return ParkingCall(sc, idxs, variant, fixTime).run(opts, nthreads, lWS, nWS)
