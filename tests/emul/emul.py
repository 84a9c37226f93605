"""ctypes loader for the DEVELOPMENT emulation of one solver CTA (tests/emul/emul.cpp).  Test tool only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libobca_emul.so")


class IpmOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double), ("mu_min", C.c_double),
                ("kappa_eps", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double), ("tau_min", C.c_double),
                ("kappa1", C.c_double), ("kappa2", C.c_double), ("kappa_sigma", C.c_double), ("s_max", C.c_double),
                ("dual_inf_tol", C.c_double), ("constr_viol_tol", C.c_double), ("compl_inf_tol", C.c_double),
                ("dw_min", C.c_double), ("dw_first", C.c_double), ("dw_max", C.c_double), ("kw_minus", C.c_double),
                ("kw_plus", C.c_double), ("kw_plus_first", C.c_double),
                ("gamma_theta", C.c_double), ("gamma_phi", C.c_double), ("delta", C.c_double), ("s_theta", C.c_double),
                ("s_phi", C.c_double), ("eta_phi", C.c_double), ("gamma_alpha", C.c_double),
                ("max_backtrack", C.c_int), ("dc", C.c_double), ("max_kick", C.c_int), ("quad_dual_ws", C.c_int)]


def build(force=False):
    srcs = [os.path.join(HERE, "emul.cpp")] + [os.path.join(HERE, "../../obca_b200/csrc", f)
                                                for f in os.listdir(os.path.join(HERE, "../../obca_b200/csrc"))
                                                if f.endswith((".cuh", ".h"))]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", SO,
                               os.path.join(HERE, "emul.cpp")])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def default_opts():
    o = IpmOpts()
    lib().emul_default_opts(C.byref(o))
    return o


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def solve_batch(sc, fixTime=0, variant="sd", opts=None, lWS=None, nWS=None, dump=False):
    """sc: dict from obca_b200.scenarios (B, N, x0, xF, rx, ry, ryaw, xWS (B,N+1,4), uWS (B,N,2), ...)."""
    B, N, nOb = sc["B"], sc["N"], sc["nOb"]
    NS = N + 1
    vOb = np.ascontiguousarray(sc["vOb"], dtype=np.int32)
    V = int(vOb.sum())
    A = np.asfortranarray(sc["A"], dtype=float); b = np.ascontiguousarray(sc["b"], dtype=float).ravel()
    x0 = np.ascontiguousarray(sc["x0"], dtype=float)
    xF = np.ascontiguousarray(np.broadcast_to(sc["xF"], (B, 4)), dtype=float)
    rx = np.ascontiguousarray(sc["rx"]); ry = np.ascontiguousarray(sc["ry"]); ryaw = np.ascontiguousarray(sc["ryaw"])
    xWS = np.ascontiguousarray(np.transpose(sc["xWS"], (0, 2, 1)))      # per problem: (N+1)x4 column-major == [4][N+1]
    uWS = np.ascontiguousarray(np.transpose(sc["uWS"], (0, 2, 1)))      # [2][N]
    lWSa = np.ascontiguousarray(np.transpose(lWS, (0, 2, 1)))           # (B, NS, V) -> [V][NS]
    nWSa = np.ascontiguousarray(np.transpose(nWS, (0, 2, 1)))
    sd = 1 if variant == "sd" else 0
    xp = np.zeros((B, NS, 4)); up = np.zeros((B, N, 2)); ts = np.zeros((B, NS))
    lp = np.zeros((B, NS, V)); npp = np.zeros((B, NS, 4 * nOb)); sl = np.zeros((B, NS, nOb))
    nd = 4 * N + 4 * nOb * NS
    duals = np.zeros((B, nd))
    status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); err = np.zeros(B); nfact = np.zeros(B, np.int32)
    W = None
    if dump:
        tot = lib().emul_layout_total(N, nOb, _p(vOb), sd)
        W = np.zeros(tot)
    f = lib().emul_parking_solve_batch
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_double] + [C.c_void_p] * 9 + \
                 [C.c_int, C.c_int, C.c_void_p] + [C.c_void_p] * 12
    rc = f(B, N, nOb, _p(vOb), _p(A), _p(b), _p(x0), _p(xF), float(sc["Ts"] if not fixTime else sc.get("Ts_fix", sc["Ts"])),
           float(sc["L"]), _p(np.ascontiguousarray(sc["ego"], dtype=float)),
           _p(np.ascontiguousarray(sc["XYbounds"], dtype=float)), _p(rx), _p(ry), _p(ryaw), _p(xWS), _p(uWS), _p(lWSa),
           _p(nWSa), int(fixTime), sd, C.cast(C.byref(opts), C.c_void_p) if opts is not None else None,
           _p(xp), _p(up), _p(ts), _p(lp), _p(npp), _p(sl), _p(duals), _p(status), _p(iters), _p(err), _p(nfact), _p(W))
    assert rc == 0, rc
    return dict(xp=xp, up=up, ts=ts, lp=lp, np=npp, sl=sl, duals=duals, status=status, iters=iters, err=err,
                nfact=nfact, W=W)


def dualmultws_batch(sc):
    B, N, nOb = sc["B"], sc["N"], sc["nOb"]
    NS = N + 1
    vOb = np.ascontiguousarray(sc["vOb"], dtype=np.int32); V = int(vOb.sum())
    A = np.asfortranarray(sc["A"], dtype=float); b = np.ascontiguousarray(sc["b"], dtype=float).ravel()
    lp = np.zeros((B, V, NS)); npp = np.zeros((B, 4 * nOb, NS)); d = np.zeros((B, nOb, NS)); its = np.zeros((B, nOb, NS), np.int32)
    f = lib().emul_dualmultws_batch
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 11
    rc = f(B, N, nOb, _p(vOb), _p(A), _p(b), _p(np.ascontiguousarray(sc["ego"], dtype=float)),
           _p(np.ascontiguousarray(sc["rx"])), _p(np.ascontiguousarray(sc["ry"])), _p(np.ascontiguousarray(sc["ryaw"])),
           _p(lp), _p(npp), _p(d), _p(its))
    assert rc == 0
    # return in (B, NS, V) "row = stage" orientation like DualMultWS.jl:81-84
    return np.transpose(lp, (0, 2, 1)).copy(), np.transpose(npp, (0, 2, 1)).copy(), np.transpose(d, (0, 2, 1)).copy(), its


def eval_batch(sc_one, arrays, fixTime, variant):
    """K1 stand-alone on the host for ONE problem.  sc_one: dict with N, nOb, vOb, A, b, x0 (4,), xF (4,), Ts, L, ego,
    XYbounds, rx, ry, ryaw (N+1,).  arrays: dict from tests/k1_maps.k1_inputs."""
    N, nOb = sc_one["N"], sc_one["nOb"]
    vOb = np.ascontiguousarray(sc_one["vOb"], dtype=np.int32); V = int(vOb.sum())
    sd = 1 if variant == "sd" else 0
    NS = N + 1
    n = 4 * NS + NS + 2 * N + V * NS + 4 * nOb * NS + (nOb * NS if sd else 0)
    m = 8 + 6 * N + 4 * nOb * NS
    c = np.zeros(m); gl = np.zeros(n); fk = np.zeros(NS)
    f = lib().emul_parking_eval_batch
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_double, C.c_double] + [C.c_void_p] * 12 + [C.c_int, C.c_int] + [C.c_void_p] * 3
    A = np.asfortranarray(sc_one["A"], dtype=float); b = np.ascontiguousarray(sc_one["b"], dtype=float).ravel()
    g = lambda a: np.ascontiguousarray(a, dtype=float) if a is not None else None
    keep = [g(sc_one["x0"]), g(sc_one["xF"]), g(sc_one["ego"]), g(sc_one["XYbounds"]), g(sc_one["rx"]), g(sc_one["ry"]), g(sc_one["ryaw"])]
    rc = f(1, N, nOb, _p(vOb), _p(A), _p(b), _p(keep[0]), _p(keep[1]), float(sc_one["Ts"]), float(sc_one["L"]), _p(keep[2]), _p(keep[3]),
           _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(arrays["xp"]), _p(arrays["up"]), _p(arrays["ts"]), _p(arrays["lp"]), _p(arrays["np"]),
           _p(arrays["sl"]), _p(arrays["y"]), int(fixTime), sd, _p(c), _p(gl), _p(fk))
    assert rc == 0
    return c, gl, fk


def quad_solve_batch(sc, variant="sd", opts=None):
    """sc: dict from obca_b200.scenarios.quadcopter_batch (B, N, Ts, R, obs (5,6), x0 (B,12), xF (B,12), xWS (B,12,N+1))."""
    B, N = sc["B"], sc["N"]; NS = N + 1
    x0 = np.ascontiguousarray(sc["x0"], dtype=float); xF = np.ascontiguousarray(sc["xF"], dtype=float)
    obs = np.ascontiguousarray(sc["obs"], dtype=float)                       # rows = obstacles -> memory 6 x 5 column-major
    xWS = np.ascontiguousarray(np.transpose(sc["xWS"], (0, 2, 1)), dtype=float)   # per problem [N+1][12] == 12 x (N+1) column-major
    sd = 1 if variant == "sd" else 0
    xp = np.zeros((B, NS, 12)); up = np.zeros((B, N, 4)); ts = np.zeros((B, NS)); lp = np.zeros((B, NS, 30)); sl = np.zeros((B, NS, 5))
    status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); err = np.zeros(B); nfact = np.zeros(B, np.int32)
    f = lib().emul_quadcopter_solve_batch
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_int] + [C.c_void_p] * 10
    rc = f(B, N, _p(x0), _p(xF), float(sc["Ts"]), float(sc["R"]), _p(obs), _p(xWS), float(sc.get("timeWS", 1.0)), sd,
           C.cast(C.byref(opts), C.c_void_p) if opts is not None else None, _p(xp), _p(up), _p(ts), _p(lp), _p(sl), _p(status),
           _p(iters), _p(err), _p(nfact))
    assert rc == 0, rc
    T = lambda a: np.transpose(a, (0, 2, 1))
    return dict(xp=T(xp), up=T(up), ts=ts, lp=T(lp), slack=T(sl), status=status, iters=iters, err=err, nfact=nfact)
