"""Host-side scenario generation (float64): river banks + goal sets -> a "scenario bank" for the GPU.

Restates the reset path of the reference -- `game_map.gen_river_poly` (ship_gym/game_map.py:22-73), the
convexification `pm.Poly` applies to each bank (models.py:180 -> cpConvexHull, tolerance 0) and
`ShipGame.gen_goal_path` (game.py:300-330) with its two fat (radius 10) horizontal segment queries per goal
(cpShapeSegmentQuery / cpPolyShapeSegmentQuery incl. the bevelled vertices).  The random draws are made in
the same order and with the same generators the reference uses (`random.gauss`, `random.randint`,
`np.random.uniform`), so `generate(..., seed=s)` yields, scenario by scenario, the maps the reference would
build on consecutive `reset()` calls after `random.seed(s); np.random.seed(s)` (tests/golden/scenarios.npz).

This module is product code: it never touches oracle/.
"""
import math
import random as _pyrandom

import numpy as np

N_GOALS = 5                # game.py:17
GOAL_QUERY_RADIUS = 10.0   # game.py:322-323
GOAL_TOLERANCE = 60.0      # game.py:316
MAX_HULL = 32              # SHIPSIM_MAX_HULL
_TINY = 2.2250738585072014e-308


# ------------------------------------------------------------------------------------------ bank vertices
def river_banks(rnd, bounds, N=10, width_frac=0.5, y_jitter=20.0, x_jitter=50.0, max_tries=1000):
    """Two raw vertex lists [left, right] in the reference's order (game_map.py:22-73).

    Per bank and per segment i=1..N: x ~ gauss(x_max, x_jitter) (the reference's `x_middle` evaluates to
    x_max, game_map.py:48), y = -100 + gauss(i*y_delta, y_jitter); redraw BOTH while x is outside
    [x_min, x_max], at most `max_tries` draws; then the two outer wall corners are appended."""
    W, H = float(bounds[0]), float(bounds[1])
    y_start = -100.0
    y_delta = (H * 1.2 - y_start) / N
    bank_width = width_frac * W / 2

    def one_bank(x_min, x_max):
        centre = x_min + (x_max - x_min)
        out = []
        for i in range(1, N + 1):
            for _ in range(max_tries):
                x = rnd.gauss(centre, x_jitter)
                y = y_start + rnd.gauss(y_delta * i, y_jitter)
                if x_min <= x <= x_max:
                    break
            out.append((x, y))
        return out

    left = one_bank(0.0, bank_width) + [(0.0, H), (0.0, 0.0)]
    right = one_bank(W - bank_width, W) + [(W, H), (W, 0.0)]
    return [left, right]


def convex_hull(points):
    """CCW convex hull, collinear points dropped, first vertex = min x then min y (cpConvexHull, tol 0)."""
    pts = sorted(set((float(x), float(y)) for x, y in points))
    if len(pts) < 3:
        return pts

    def chain(seq):
        out = []
        for p in seq:
            while len(out) >= 2:
                (ox, oy), (ax, ay) = out[-2], out[-1]
                if (ax - ox) * (p[1] - oy) - (ay - oy) * (p[0] - ox) <= 0.0:
                    out.pop()
                else:
                    break
            out.append(p)
        return out

    lo = chain(pts)
    hi = chain(pts[::-1])
    return lo[:-1] + hi[:-1]


# ------------------------------------------------------------------------------------------ fat segment query
def _planes(hull):
    n = len(hull)
    out = []
    for i in range(n):
        ax, ay = hull[i - 1]
        bx, by = hull[i]
        ex, ey = bx - ax, by - ay
        ln = math.hypot(ex, ey)
        out.append((ey / ln, -ex / ln))
    return out


def _point_distance(hull, normals, px, py):
    """Signed distance to a convex polygon, negative inside (cpPolyShapePointQuery)."""
    outside = False
    best = math.inf
    n = len(hull)
    for i in range(n):
        ax, ay = hull[i - 1]
        bx, by = hull[i]
        nx, ny = normals[i]
        if nx * (px - bx) + ny * (py - by) > 0.0:
            outside = True
        ex, ey = ax - bx, ay - by
        t = (ex * (px - bx) + ey * (py - by)) / (ex * ex + ey * ey)
        t = min(max(t, 0.0), 1.0)
        d = math.hypot(px - (bx + ex * t), py - (by + ey * t))
        best = min(best, d)
    return best if outside else -best


def fat_segment_hit_x(hull, a, b, radius):
    """x coordinate of `cpShapeSegmentQuery(hull, a, b, radius).point`, or None when the query misses."""
    normals = _planes(hull)
    ax, ay = a
    bx, by = b
    if _point_distance(hull, normals, ax, ay) <= radius:
        return bx                                   # alpha = 0: `point` is left at the segment end
    n = len(hull)
    hit_x, alpha = None, 1.0
    for i in range(n):
        nx, ny = normals[i]
        vx, vy = hull[i]
        an = ax * nx + ay * ny
        d = an - (vx * nx + vy * ny) - radius
        if d < 0.0:
            continue
        bn = bx * nx + by * ny
        t = d / max(an - bn, _TINY)
        if t < 0.0 or t > 1.0:
            continue
        qx, qy = ax + (bx - ax) * t, ay + (by - ay) * t
        along = nx * qy - ny * qx
        ux, uy = hull[i - 1]
        if nx * uy - ny * ux <= along <= nx * vy - ny * vx:
            hit_x, alpha = qx - nx * radius, t
    for (cx, cy) in hull:                            # bevelled vertices (CircleSegmentQuery)
        dax, day, dbx, dby = ax - cx, ay - cy, bx - cx, by - cy
        daa, dab, dbb = dax * dax + day * day, dax * dbx + day * dby, dbx * dbx + dby * dby
        qa = daa - 2.0 * dab + dbb
        qb = dab - daa
        det = qb * qb - qa * (daa - radius * radius)
        if det >= 0.0 and qa != 0.0:
            t = (-qb - math.sqrt(det)) / qa
            if 0.0 <= t <= 1.0 and t < alpha:
                mx, my = dax + (dbx - dax) * t, day + (dby - day) * t
                ln = math.hypot(mx, my)
                hit_x, alpha = ax + (bx - ax) * t - (mx / ln) * radius, t
    return hit_x


def _first_bank_hit(hulls, a, b, radius):
    """`space.segment_query(a, b, radius, filter)[0].point.x` over the static bank shapes in insertion order
    (game.py:322-323).  The spatial index only offers shapes whose box the THIN segment crosses."""
    lo_x, hi_x = min(a[0], b[0]), max(a[0], b[0])
    for hull in hulls:
        xs = [v[0] for v in hull]
        ys = [v[1] for v in hull]
        if a[1] < min(ys) or a[1] > max(ys) or hi_x < min(xs) or lo_x > max(xs):
            continue
        x = fat_segment_hit_x(hull, a, b, radius)
        if x is not None:
            return x
    return None


def goal_path(rnd, nprnd, hulls, bounds, n=N_GOALS):
    """Goal centres in creation order (game.py:300-330)."""
    W, H = float(bounds[0]), float(bounds[1])
    y_delta = H / (n + 1)
    goals = []
    for i in range(1, n + 1):
        y = y_delta * i + rnd.randint(-20, 20)
        left = _first_bank_hit(hulls, (W / 2, y), (0.0, y), GOAL_QUERY_RADIUS)
        right = _first_bank_hit(hulls, (W / 2, y), (W, y), GOAL_QUERY_RADIUS)
        if left is None or right is None:            # the reference's `except` branch (game.py:328-330)
            x = (W / 2) * i + rnd.randint(-50, 50)
        else:
            x = nprnd.uniform(left + GOAL_TOLERANCE, right - GOAL_TOLERANCE)
        goals.append((x, y))
    return goals


# ------------------------------------------------------------------------------------------ the bank
class ScenarioBank(object):
    """`n` scenarios as float64 arrays: hull_xy [n,2,maxv,2] (CCW, zero padded), hull_n [n,2], goals [n,5,2];
    `raw` keeps the un-convexified vertex lists (for inspection / tests)."""

    def __init__(self, hull_xy, hull_n, goals, bounds, raw=None):
        self.hull_xy = np.ascontiguousarray(hull_xy, dtype=np.float64)
        self.hull_n = np.ascontiguousarray(hull_n, dtype=np.int32)
        self.goals = np.ascontiguousarray(goals, dtype=np.float64)
        self.bounds = (float(bounds[0]), float(bounds[1]))
        self.raw = raw

    def __len__(self):
        return self.hull_xy.shape[0]

    @property
    def maxv(self):
        return self.hull_xy.shape[2]

    def as_dict(self):
        return dict(hull_xy=self.hull_xy, hull_n=self.hull_n, goals=self.goals)

    @classmethod
    def from_hulls(cls, hulls_list, goals_list, bounds):
        n = len(hulls_list)
        maxv = max(len(h) for hs in hulls_list for h in hs)
        if maxv > MAX_HULL:
            raise ValueError("a bank hull has %d vertices; at most %d are supported" % (maxv, MAX_HULL))
        hull_xy = np.zeros((n, 2, maxv, 2))
        hull_n = np.zeros((n, 2), dtype=np.int32)
        for s, hs in enumerate(hulls_list):
            for b, h in enumerate(hs):
                hull_xy[s, b, :len(h)] = np.asarray(h, dtype=np.float64)
                hull_n[s, b] = len(h)
        return cls(hull_xy, hull_n, np.asarray(goals_list, dtype=np.float64).reshape(n, N_GOALS, 2), bounds)

    @classmethod
    def generate(cls, n, bounds=(600, 600), seed=0, map_N=10, width_frac=0.5):
        """`n` consecutive reference resets worth of maps after seeding python `random` and numpy with `seed`."""
        rnd = _pyrandom.Random(seed)
        nprnd = np.random.RandomState(seed)
        hulls_list, goals_list, raws = [], [], []
        for _ in range(n):
            raw = river_banks(rnd, bounds, N=map_N, width_frac=width_frac)
            hulls = [convex_hull(v) for v in raw]
            goals = goal_path(rnd, nprnd, hulls, bounds)
            hulls_list.append(hulls)
            goals_list.append(goals)
            raws.append(raw)
        bank = cls.from_hulls(hulls_list, goals_list, bounds)
        bank.raw = raws
        return bank
