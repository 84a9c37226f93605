"""Velocity-profile smoothing of the reference's warm-start pipeline (AutonomousParking/veloSmooth.jl:29-109): the piecewise
constant +-v0 speed profile read off the Hybrid A* path gets linear ramps of |v0| / amax seconds at every start, stop and
direction change; returns the smoothed speed and its finite-difference acceleration."""
from __future__ import annotations

import numpy as np

from .grid_policy import jround

PAD = 19          # the reference embeds v at offset 19 of arrays 40 longer (veloSmooth.jl:30-41)


def velo_smooth(v, amax, Ts):
    v = np.asarray(v, float).ravel()
    n = v.size
    v_ex = np.zeros(n + 40)
    v_bar = np.zeros((4, n + 40))
    v_ex[PAD:PAD + n] = v
    v_bar[:, PAD:PAD + n] = v
    v0 = abs(v[0])
    cut1, cut2 = 0.25 * v0, 1.25 * v0
    acc = jround(v0 / amax / Ts)
    dv = np.diff(v_ex)
    # 1-based indices as in the reference (find(...) on diff(v_ex))
    idx1 = list(np.flatnonzero((dv > cut1) & (dv < cut2)) + 1)
    idx2 = list(np.flatnonzero(dv > cut2) + 1)
    idx3 = list(np.flatnonzero((dv < -cut1) & (dv > -cut2)) + 1)
    idx4 = list(np.flatnonzero(dv < -cut2) + 1)
    if idx1 and idx1[0] == 19:
        idx1[0] += 1
    if idx3 and idx3[0] == 19:
        idx3[0] += 1
    ex = lambda i: v_ex[i - 1]                      # 1-based read

    def put(row, lo, hi, a, b):                     # v_bar[row, lo:hi] = linspace(a, b, hi - lo + 1), 1-based inclusive
        v_bar[row, lo - 1:hi] = np.linspace(a, b, hi - lo + 1)

    for i in idx1:
        if ex(i) > cut1 or ex(i + 1) > cut1:
            put(0, i, i + acc, 0.0, v0)
        elif ex(i) < -cut1 or ex(i + 1) < -cut1:
            put(0, i - acc + 1, i + 1, -v0, 0.0)
    for i in idx3:
        if ex(i) > cut1 or ex(i + 1) > cut1:
            put(1, i - acc + 1, i + 1, v0, 0.0)
        elif ex(i) < -cut1 or ex(i + 1) < -cut1:
            put(1, i, i + acc, 0.0, -v0)
    for i in idx2:
        put(2, i - acc, i + acc, -v0, v0)
    for i in idx4:
        put(3, i - acc, i + acc, v0, -v0)
    sl = slice(PAD, PAD + n)
    vb = v_bar[:, sl]; ve = v_ex[sl]
    v_bar2 = np.where(vb == 0.0, vb, np.where(np.sign(ve)[None, :] != np.sign(vb), ve[None, :], vb))
    v_mm = np.where(ve > 0, v_bar2.min(axis=0), v_bar2.max(axis=0))
    a = np.diff(v_mm) / Ts
    return v_mm, a
