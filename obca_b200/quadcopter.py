"""Host-side mirror of the reference's quadcopter NLP drivers, same names / argument order / return tuples:

    QuadcopterSignedDist  QuadcopterNavigation/QuadcopterSignedDist.jl:25 -> (xp, up, timeScalep, exitflag, time, lp, status) :298
    QuadcopterDist        QuadcopterNavigation/QuadcopterDist.jl:25       -> same 7-tuple :280
    constrSatisfaction    QuadcopterNavigation/constrSatisfaction.jl:25   -> Bool

All computation happens in libobca.so on the GPU (obca_quadcopter_solve_batch / obca_check_quadcopter)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f64, lib, ptr


def quadcopter_solve_batch(x0, xF, N, Ts, R, obs, xWS, timeWS=1.0, signed_dist=1, opts=None):
    """x0, xF (B,12); obs (5,6) = ob1..ob5; xWS (B,12,N+1).  Returns dict with Julia-shaped stacks:
    xp (B,12,N+1), up (B,4,N), ts (B,N+1), lp (B,30,N+1), slack (B,5,N+1), exitflag, iters, kkt_err, time."""
    x0 = f64(np.atleast_2d(x0)); B = x0.shape[0]; NS = N + 1
    xF = f64(np.broadcast_to(np.asarray(xF, float).reshape(-1, 12), (B, 12)))
    ob = f64(np.asarray(obs, float).reshape(5, 6))                       # rows = obstacles == 6 x 5 column-major
    xw = f64(np.transpose(np.asarray(xWS, float).reshape(B, 12, -1)[:, :, :NS], (0, 2, 1)))
    xp = np.zeros((B, NS, 12)); up = np.zeros((B, N, 4)); ts = np.zeros((B, NS)); lp = np.zeros((B, NS, 30)); sl = np.zeros((B, NS, 5))
    ef = np.zeros(B, np.int32); it = np.zeros(B, np.int32); err = np.zeros(B); sec = np.zeros(1)
    if opts is None:
        opts = _lib.default_opts(); opts.max_iter = 3000                 # the reference leaves Ipopt's default max_iter
    check(lib().obca_quadcopter_solve_batch(C.c_int(B), C.c_int(N), ptr(x0), ptr(xF), C.c_double(Ts), C.c_double(R), ptr(ob), ptr(xw),
                                            C.c_double(float(timeWS)), C.c_int(int(signed_dist)), C.byref(opts), ptr(xp), ptr(up), ptr(ts),
                                            ptr(lp), ptr(sl), ptr(ef), ptr(it), ptr(err), ptr(sec)))
    T = lambda a: np.transpose(a, (0, 2, 1))
    return dict(xp=T(xp), up=T(up), ts=ts, lp=T(lp), slack=T(sl), exitflag=ef, iters=it, kkt_err=err, time=float(sec[0]))


def _single(sd, x0, xF, N, Ts, R, ob1, ob2, ob3, ob4, ob5, xWS, uWS, timeWS):
    obs = np.stack([np.asarray(o, float).ravel() for o in (ob1, ob2, ob3, ob4, ob5)])
    r = quadcopter_solve_batch(np.asarray(x0, float).reshape(1, 12), np.asarray(xF, float).reshape(1, 12), N, Ts, R, obs,
                               np.asarray(xWS, float)[None], timeWS, sd)
    ef = int(r["exitflag"][0])
    return r["xp"][0], r["up"][0], r["ts"][0], ef, r["time"], r["lp"][0], ("Optimal" if ef >= 1 else "Error")


def QuadcopterSignedDist(x0, xF, N, Ts, R, ob1, ob2, ob3, ob4, ob5, xWS, uWS, timeWS):
    return _single(1, x0, xF, N, Ts, R, ob1, ob2, ob3, ob4, ob5, xWS, uWS, timeWS)


def QuadcopterDist(x0, xF, N, Ts, R, ob1, ob2, ob3, ob4, ob5, xWS, uWS, timeWS):
    return _single(0, x0, xF, N, Ts, R, ob1, ob2, ob3, ob4, ob5, xWS, uWS, timeWS)


def check_quadcopter_batch(x, u, timeScale, x0, xF, Ts, lam, obs, R, opts=None):
    """x (B,12,N+1), u (B,4,N), timeScale (B,N+1), lam (B,30,N+1)."""
    x = np.asarray(x, float); B = x.shape[0]; N = x.shape[2] - 1
    T = lambda a: f64(np.transpose(np.asarray(a, float), (0, 2, 1)))
    x0 = f64(np.broadcast_to(np.asarray(x0, float).reshape(-1, 12), (B, 12))); xF = f64(np.broadcast_to(np.asarray(xF, float).reshape(-1, 12), (B, 12)))
    xx, uu, ll = T(x), T(u), T(lam)
    tt = f64(np.asarray(timeScale, float).reshape(B, N + 1))
    ob = f64(np.asarray(obs, float).reshape(5, 6))
    feas = np.zeros(B, np.int32); worst = np.zeros(B)
    o = opts if opts is not None else _lib.default_opts()
    check(lib().obca_check_quadcopter(C.c_int(B), C.c_int(N), ptr(xx), ptr(uu), ptr(tt), ptr(x0), ptr(xF), C.c_double(Ts), ptr(ll), ptr(ob),
                                      C.c_double(R), C.byref(o), ptr(feas), ptr(worst)))
    return feas, worst


def constrSatisfaction(x, u, timeScale, x0, xF, Ts, lam, ob1, ob2, ob3, ob4, ob5, R):
    obs = np.stack([np.asarray(o, float).ravel() for o in (ob1, ob2, ob3, ob4, ob5)])
    feas, _ = check_quadcopter_batch(np.asarray(x, float)[None], np.asarray(u, float)[None], np.asarray(timeScale, float).reshape(1, -1),
                                     x0, xF, Ts, np.asarray(lam, float)[None], obs, R)
    return bool(feas[0])
