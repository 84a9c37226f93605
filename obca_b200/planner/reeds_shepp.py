"""Reeds-Shepp curves: host-side restatement of AutonomousParking/reeds_shepp.jl (the analytic expansion and heuristic of the
reference's Hybrid A* warm-start generator).  Not on the GPU path: it produces the (rx, ry, ryaw)
path the NLP drivers are warm-started with.

Everything is computed in the normalised frame of the reference (unit turning radius: lengths are angles for arcs and
distance x max-curvature for straights, `generate_path`, reeds_shepp.jl:782-800) and scaled back by 1/maxc.

Structure: instead of one hand-expanded function per path family (reeds_shepp.jl:233-670) the candidate words are generated
from  (base word solver, segment types, how (t, u, v) map to segment lengths)  x  the symmetries of the problem:
    time flip  (x, y, phi) -> (-x,  y, -phi): all lengths change sign
    reflection (x, y, phi) -> ( x, -y, -phi): L <-> R
    backwards  (x, y, phi) -> (xb, yb, phi), xb = x cos(phi) + y sin(phi), yb = x sin(phi) - y cos(phi): word reversed
in the reference's order, because two reference quirks make the ORDER observable: `set_path` drops a candidate when an
earlier one of the same type has  sum(old.lengths - new.lengths) <= 0.01  (signed sum, reeds_shepp.jl:212-219), and
`calc_shortest_path` keeps the LAST of equally short paths (`<=`, :67-72).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

STEP_SIZE = 0.1                       # reeds_shepp.jl:35


@dataclass
class Path:                            # reeds_shepp.jl:37-45
    lengths: List[float]               # signed length of every segment (+ forward, - backward)
    ctypes: List[str]                  # "S" | "L" | "R" per segment
    L: float = 0.0                     # total length
    x: List[float] = field(default_factory=list)
    y: List[float] = field(default_factory=list)
    yaw: List[float] = field(default_factory=list)
    directions: List[int] = field(default_factory=list)


def pi_2_pi(a: float) -> float:        # reeds_shepp.jl:47-56
    while a > math.pi:
        a -= 2.0 * math.pi
    while a < -math.pi:
        a += 2.0 * math.pi
    return a


def polar(x: float, y: float) -> Tuple[float, float]:
    return math.sqrt(x * x + y * y), math.atan2(y, x)      # (not math.hypot: the native planner computes the same expression)


def mod2pi(x: float) -> float:         # reeds_shepp.jl:146-156 (Julia mod: result has the sign of the divisor, i.e. in [0, 2 pi))
    v = x % (2.0 * math.pi)
    if v < -math.pi:
        v += 2.0 * math.pi
    elif v > math.pi:
        v -= 2.0 * math.pi
    return v


Word = Optional[Tuple[float, float, float]]


# ---- base words: each returns (t, u, v) or None (reeds_shepp.jl:159-204, 249-269, 371-424, 481-516, 624-642) ----
def LSL(x, y, phi) -> Word:
    u, t = polar(x - math.sin(phi), y - 1.0 + math.cos(phi))
    if t >= 0.0:
        v = mod2pi(phi - t)
        if v >= 0.0:
            return t, u, v
    return None


def LSR(x, y, phi) -> Word:
    u1, t1 = polar(x + math.sin(phi), y - 1.0 - math.cos(phi))
    u1 = u1 * u1
    if u1 >= 4.0:
        u = math.sqrt(u1 - 4.0)
        theta = math.atan2(2.0, u)
        t = mod2pi(t1 + theta)
        v = mod2pi(t - phi)
        if t >= 0.0 and v >= 0.0:
            return t, u, v
    return None


def LRL(x, y, phi) -> Word:
    u1, t1 = polar(x - math.sin(phi), y - 1.0 + math.cos(phi))
    if u1 <= 4.0:
        u = -2.0 * math.asin(0.25 * u1)
        t = mod2pi(t1 + 0.5 * u + math.pi)
        v = mod2pi(phi - t + u)
        if t >= 0.0 and u <= 0.0:
            return t, u, v
    return None


def SLS(x, y, phi) -> Word:
    phi = mod2pi(phi)
    if phi > 0.0 and phi < math.pi * 0.99 and y != 0.0:
        xd = -y / math.tan(phi) + x
        t = xd - math.tan(phi / 2.0)
        u = phi
        r = math.sqrt((x - xd) ** 2 + y * y)
        v = (r if y > 0.0 else -r) - math.tan(phi / 2.0)
        return t, u, v
    return None


def _tau_omega(u, v, xi, eta, phi):
    delta = mod2pi(u - v)
    A = math.sin(u) - math.sin(delta)
    B = math.cos(u) - math.cos(delta) - 1.0
    t1 = math.atan2(eta * A - xi * B, xi * A + eta * B)
    t2 = 2.0 * (math.cos(delta) - math.cos(v) - math.cos(u)) + 3.0
    tau = mod2pi(t1 + math.pi) if t2 < 0 else mod2pi(t1)
    omega = mod2pi(tau - u + v - phi)
    return tau, omega


def LRLRn(x, y, phi) -> Word:
    xi = x + math.sin(phi); eta = y - 1.0 - math.cos(phi)
    rho = 0.25 * (2.0 + math.sqrt(xi * xi + eta * eta))
    if rho <= 1.0:
        u = math.acos(rho)
        t, v = _tau_omega(u, -u, xi, eta, phi)
        if t >= 0.0 and v <= 0.0:
            return t, u, v
    return None


def LRLRp(x, y, phi) -> Word:
    xi = x + math.sin(phi); eta = y - 1.0 - math.cos(phi)
    rho = (20.0 - xi * xi - eta * eta) / 16.0
    if 0.0 <= rho <= 1.0:
        u = -math.acos(rho)
        if u >= -0.5 * math.pi:
            t, v = _tau_omega(u, u, xi, eta, phi)
            if t >= 0.0 and v >= 0.0:
                return t, u, v
    return None


def LRSR(x, y, phi) -> Word:
    xi = x + math.sin(phi); eta = y - 1.0 - math.cos(phi)
    rho, theta = polar(-eta, xi)
    if rho >= 2.0:
        t = theta
        u = 2.0 - rho
        v = mod2pi(t + 0.5 * math.pi - phi)
        if t >= 0.0 and u <= 0.0 and v <= 0.0:
            return t, u, v
    return None


def LRSL(x, y, phi) -> Word:
    xi = x - math.sin(phi); eta = y - 1.0 + math.cos(phi)
    rho, theta = polar(xi, eta)
    if rho >= 2.0:
        r = math.sqrt(rho * rho - 4.0)
        u = 2.0 - r
        t = mod2pi(theta + math.atan2(r, -2.0))
        v = mod2pi(phi - 0.5 * math.pi - t)
        if t >= 0.0 and u <= 0.0 and v <= 0.0:
            return t, u, v
    return None


def LRSLR(x, y, phi) -> Word:
    xi = x + math.sin(phi); eta = y - 1.0 - math.cos(phi)
    rho, _ = polar(xi, eta)
    if rho >= 2.0:
        u = 4.0 - math.sqrt(rho * rho - 4.0)
        if u <= 0.0:
            t = mod2pi(math.atan2((4.0 - u) * xi - 2.0 * eta, -2.0 * xi + (u - 4.0) * eta))
            v = mod2pi(t - phi)
            if t >= 0.0 and v >= 0.0:
                return t, u, v
    return None


HP = 0.5 * math.pi
_SWAP = {"L": "R", "R": "L", "S": "S"}


# Reference quirk (reeds_shepp.jl:212-219): a candidate is dropped when an earlier one of the same segment types has
# sum(old.lengths - new.lengths) <= 0.01 -- a SIGNED sum, so any earlier word whose lengths add up to less blocks the new one even
# if the new one is shorter.  Measured here: for 21 % of random pose pairs the "shortest" path is then not the shortest (up to
# several metres longer) and L(a -> b) != L(b -> a).  REFERENCE_DEDUP = True reproduces the reference (default: same warm
# starts as main.jl); False compares |differences| per segment, which keeps every distinct word (true Reeds-Shepp distances).
REFERENCE_DEDUP = True


def _set_path(paths: List[Path], lengths: List[float], ctypes: List[str]) -> None:
    """reeds_shepp.jl:207-230."""
    for tp in paths:
        if tp.ctypes != ctypes:
            continue
        if REFERENCE_DEDUP:
            if sum(a - b for a, b in zip(tp.lengths, lengths)) <= 0.01:
                return
        elif sum(abs(a - b) for a, b in zip(tp.lengths, lengths)) <= 0.01:
            return
    L = sum(abs(l) for l in lengths)
    if L < 0.01:                       # the reference asserts L >= 0.01 (Base.Test.@test, :226); a zero-length word is no path
        return
    paths.append(Path(list(lengths), list(ctypes), L))


def _four_symmetries(paths, solver, word, pattern, x, y, phi, reverse=False):
    """identity, time flip, reflection, both -- in the reference's order."""
    for fx, fy, fphi, flip, reflect in ((x, y, phi, False, False), (-x, y, -phi, True, False),
                                        (x, -y, -phi, False, True), (-x, -y, phi, True, True)):
        w = solver(fx, fy, fphi)
        if w is None:
            continue
        lengths = pattern(*w)
        types = list(word)
        if reverse:
            lengths = lengths[::-1]; types = types[::-1]
        if flip:
            lengths = [-l for l in lengths]
        if reflect:
            types = [_SWAP[c] for c in types]
        _set_path(paths, lengths, types)


def generate_path(q0, q1, maxc: float) -> List[Path]:
    """All candidate words from q0 to q1 in the normalised frame (reeds_shepp.jl:782-800)."""
    dx = q1[0] - q0[0]; dy = q1[1] - q0[1]; dth = q1[2] - q0[2]
    c = math.cos(q0[2]); s = math.sin(q0[2])
    x = (c * dx + s * dy) * maxc
    y = (-s * dx + c * dy) * maxc
    phi = dth
    paths: List[Path] = []
    # SCS (:233-246): straight-arc-straight, identity and reflection only
    w = SLS(x, y, phi)
    if w is not None:
        _set_path(paths, list(w), ["S", "L", "S"])
    w = SLS(x, -y, -phi)
    if w is not None:
        _set_path(paths, list(w), ["S", "R", "S"])
    xb = x * math.cos(phi) + y * math.sin(phi)
    yb = x * math.sin(phi) - y * math.cos(phi)
    tuv = lambda t, u, v: [t, u, v]
    ccsc = lambda t, u, v: [t, -HP, u, v]
    _four_symmetries(paths, LSL, "LSL", tuv, x, y, phi)                                   # CSC    (:272-315)
    _four_symmetries(paths, LSR, "LSR", tuv, x, y, phi)
    _four_symmetries(paths, LRL, "LRL", tuv, x, y, phi)                                   # CCC    (:318-368)
    _four_symmetries(paths, LRL, "LRL", tuv, xb, yb, phi, reverse=True)
    _four_symmetries(paths, LRLRn, "LRLR", lambda t, u, v: [t, u, -u, v], x, y, phi)      # CCCC   (:427-478)
    _four_symmetries(paths, LRLRp, "LRLR", lambda t, u, v: [t, u, u, v], x, y, phi)
    _four_symmetries(paths, LRSL, "LRSL", ccsc, x, y, phi)                                # CCSC   (:519-621)
    _four_symmetries(paths, LRSR, "LRSR", ccsc, x, y, phi)
    _four_symmetries(paths, LRSL, "LRSL", ccsc, xb, yb, phi, reverse=True)
    _four_symmetries(paths, LRSR, "LRSR", ccsc, xb, yb, phi, reverse=True)
    _four_symmetries(paths, LRSLR, "LRSLR", lambda t, u, v: [t, -HP, u, -HP, v], x, y, phi)   # CCSCC  (:645-670)
    return paths


# ---- sampling a word (reeds_shepp.jl:673-779) --------------------------------------------------------------------------
def _interpolate(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw, directions):
    if m == "S":
        px[ind] = ox + l / maxc * math.cos(oyaw)
        py[ind] = oy + l / maxc * math.sin(oyaw)
        pyaw[ind] = oyaw
    else:
        ldx = math.sin(l) / maxc
        ldy = (1.0 - math.cos(l)) / (maxc if m == "L" else -maxc)
        gdx = math.cos(-oyaw) * ldx + math.sin(-oyaw) * ldy
        gdy = -math.sin(-oyaw) * ldx + math.cos(-oyaw) * ldy
        px[ind] = ox + gdx
        py[ind] = oy + gdy
        pyaw[ind] = oyaw + l if m == "L" else oyaw - l
    directions[ind] = 1 if l > 0.0 else -1


def generate_local_course(L, lengths, mode, maxc, step_size):
    """Sample the word every `step_size` (normalised) plus the segment end points.  0-based restatement of :673-741, including
    the trailing-zero trimming (a course that ends exactly at x = 0 loses its last points, as in the reference)."""
    npoint = int(L / step_size) + 2 * len(lengths) + 8      # reference: + length(lengths) + 3 (:678), which can overflow by a point
                                                            # when a segment length is a whole number of steps up to round-off
    px = [0.0] * npoint; py = [0.0] * npoint; pyaw = [0.0] * npoint; directions = [0] * npoint
    ind = 1
    directions[0] = 1 if lengths[0] > 0.0 else -1
    d = step_size if lengths[0] > 0.0 else -step_size
    pd = d
    ll = 0.0
    for i, (m, l) in enumerate(zip(mode, lengths)):
        d = step_size if l > 0.0 else -step_size
        ox, oy, oyaw = px[ind], py[ind], pyaw[ind]
        ind -= 1
        if i >= 1 and lengths[i - 1] * lengths[i] > 0:
            pd = -d - ll
        else:
            pd = d - ll
        while abs(pd) <= abs(l):
            ind += 1
            _interpolate(ind, pd, m, maxc, ox, oy, oyaw, px, py, pyaw, directions)
            pd += d
        ll = l - pd - d
        ind += 1
        _interpolate(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw, directions)
    while px and px[-1] == 0.0:
        px.pop(); py.pop(); pyaw.pop(); directions.pop()
    return px, py, pyaw, directions


def calc_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE) -> List[Path]:
    """reeds_shepp.jl:99-120: every candidate, sampled and transformed to the world frame."""
    q0 = (sx, sy, syaw); q1 = (gx, gy, gyaw)
    paths = generate_path(q0, q1, maxc)
    c = math.cos(-q0[2]); s = math.sin(-q0[2])
    for p in paths:
        x, y, yaw, directions = generate_local_course(p.L, p.lengths, p.ctypes, maxc, step_size * maxc)
        p.x = [c * ix + s * iy + q0[0] for ix, iy in zip(x, y)]
        p.y = [-s * ix + c * iy + q0[1] for ix, iy in zip(x, y)]
        p.yaw = [pi_2_pi(iyaw + q0[2]) for iyaw in yaw]
        p.directions = directions
        p.lengths = [l / maxc for l in p.lengths]
        p.L = p.L / maxc
    return paths


def calc_shortest_path(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE) -> Optional[Path]:
    """reeds_shepp.jl:59-76 (the last of equally short candidates wins)."""
    paths = calc_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size)
    best = None
    minL = math.inf
    for p in paths:
        if p.L <= minL:
            minL = p.L; best = p
    return best


def calc_shortest_path_length(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE) -> float:
    """reeds_shepp.jl:79-96."""
    paths = generate_path((sx, sy, syaw), (gx, gy, gyaw), maxc)
    return min((p.L / maxc for p in paths), default=math.inf)
