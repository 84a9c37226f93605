"""Vehicle / obstacle-point collision test of the reference's Hybrid A* (AutonomousParking/collision_check.jl:40-98): a bubble
around the car centre pre-selects nearby obstacle points (KD-tree), then each of them is tested against the car rectangle
with the winding-angle rule (sum of the signed angles subtended by the four edges >= pi  <=>  the point is inside)."""
from __future__ import annotations

import math

import numpy as np

B = 1.0        # [m] rear axle -> back end            (collision_check.jl:31)
C = 3.7        # [m] rear axle -> front end           (:32)
I = 2.0        # [m] width                            (:33)
WBUBBLE_DIST = (B + C) / 2.0 - B     # rear axle -> centre of the covering bubble (:34)
WBUBBLE_R = (B + C) / 2.0            # bubble radius (:35)
_VRX = np.array([C, C, -B, -B, C])
_VRY = np.array([-I / 2.0, I / 2.0, I / 2.0, -I / 2.0, -I / 2.0])


def rect_check(ix, iy, iyaw, ox, oy) -> bool:
    """True if none of the points (ox, oy) lies inside the car rectangle at pose (ix, iy, iyaw)  (collision_check.jl:58-98)."""
    ox = np.asarray(ox, float); oy = np.asarray(oy, float)
    if ox.size == 0:
        return True
    c, s = math.cos(-iyaw), math.sin(-iyaw)
    tx, ty = ox - ix, oy - iy
    lx = c * tx - s * ty
    ly = s * tx + c * ty
    x1 = _VRX[None, :-1] - lx[:, None]; y1 = _VRY[None, :-1] - ly[:, None]
    x2 = _VRX[None, 1:] - lx[:, None]; y2 = _VRY[None, 1:] - ly[:, None]
    d1 = np.hypot(x1, y1); d2 = np.hypot(x2, y2)
    th1 = np.arctan2(y1, x1)
    tty = -np.sin(th1) * x2 + np.cos(th1) * y2
    with np.errstate(invalid="ignore", divide="ignore"):
        tmp = np.minimum((x1 * x2 + y1 * y2) / (d1 * d2), 1.0)
        ang = np.arccos(tmp)
    sumangle = np.where(tty >= 0.0, ang, -ang).sum(axis=1)
    return not bool((sumangle >= math.pi).any())


def check_collision(x, y, yaw, kdtree, ox, oy) -> bool:
    """True = collision free (collision_check.jl:40-55).  kdtree: scipy cKDTree over the obstacle points (ox, oy)."""
    ox = np.asarray(ox, float); oy = np.asarray(oy, float)
    for ix, iy, iyaw in zip(x, y, yaw):
        cx = ix + WBUBBLE_DIST * math.cos(iyaw)
        cy = iy + WBUBBLE_DIST * math.sin(iyaw)
        ids = kdtree.query_ball_point([cx, cy], WBUBBLE_R)
        if not ids:
            continue
        if not rect_check(ix, iy, iyaw, ox[ids], oy[ids]):
            return False
    return True
