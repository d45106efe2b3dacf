"""minimunk -- a float64, pure-Python restatement of the slice of pymunk 5.4.0 / Chipmunk2D 7.0.2
that CapAI/ship-sim-gym touches.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ship_sim_gym_b200/) may import this.

Why it exists: the reference (`/root/reference/ship_gym/*.py`) delegates all of its arithmetic to
`libchipmunk.so` through pymunk's cffi layer (requirements.txt:78 pins pymunk==5.4.0, which bundles
Chipmunk2D 7.0.2).  Neither pymunk nor Chipmunk is installable in this image (no network, no wheel,
no source on disk), so the reference cannot run as shipped.  This module provides a module called
`pymunk` exposing exactly the API surface the reference imports, so that the *unmodified* reference
Python (ShipGame / ShipEnv / LiDAR / Ship / PolyEnv / gen_river_poly) can be imported from
/root/reference and executed here to generate golden vectors (oracle/make_golden.py).  What that pins:
every line of the reference's own logic (action decode, lidar fan, sticky vals, reward/done quirks,
obs packing/history, goal placement, RNG call order).  What it does NOT pin: the Chipmunk arithmetic
below, which is restated from the published Chipmunk2D 7.0.2 algorithm (source files named at each
function) and is unverified against a live libchipmunk -- "Chipmunk layer: parity unpinned".

API surface covered (call sites in the reference):
  Vec2d                                    models.py:54,65,70,93,107,146  game.py:75,274,340-343
  Body(mass, moment) / Body(None, None, body_type=Body.STATIC)           models.py:92,195 game.py:84
  Body.position / angle / velocity / angular_velocity / center_of_gravity
  Body.apply_force_at_local_point          models.py:130,133   (cpBody.c cpBodyApplyForceAtLocalPoint)
  Poly(body, vertices)  .bb .segment_query .collision_type .friction .color   models.py:96-100,180
  Circle(body, radius, offset)             game.py:87
  moment_for_poly / moment_for_circle      models.py:89  game.py:83   (chipmunk.c)
  ShapeFilter(categories=, mask=), ShapeFilter.ALL_MASKS                 models.py:102 game.py:317
  Space(): damping, add, remove, step, add_collision_handler(.begin), segment_query, debug_draw
                                            game.py:71,89,113,194,252,269-270,292-298,322-323
  Transform.identity(), BB.center()/merge  models.py:109,200-203
  SpaceDebugDrawOptions.DRAW_SHAPES        game.py:203 (debug drawing is a no-op here)

Not restated (documented gap): the contact impulse solver (cpArbiterApplyImpulse, 10 iterations).  It
only changes the ship's velocity on/after the step in which `colliding` becomes True, i.e. the step
on which the reference episode ends (ship_env.py:120-121); it is unobservable before the reset.
Goal bodies (dynamic, m=1) therefore never move here, as in the reference away from bank contact.
"""
import math

version = "5.4.0-minimunk"
chipmunk_version = "7.0.2-restated"
inf = float("inf")
DBL_MIN = 2.2250738585072014e-308


# --------------------------------------------------------------------------------------------- Vec2d
class Vec2d(object):
    """Mutable 2-vector (pymunk<6 Vec2d is a mutable list-like; the reference relies on that at
    models.py:146 `self.point_of_thrust.x = ...`)."""
    __slots__ = ("x", "y")

    def __init__(self, x_or_pair=None, y=None):
        if x_or_pair is None:
            self.x, self.y = 0.0, 0.0
        elif y is None:
            self.x, self.y = x_or_pair[0], x_or_pair[1]
        else:
            self.x, self.y = x_or_pair, y

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __setitem__(self, i, v):
        if i == 0:
            self.x = v
        elif i == 1:
            self.y = v
        else:
            raise IndexError(i)

    def __iter__(self):
        yield self.x
        yield self.y

    def __repr__(self):
        return "Vec2d(%r, %r)" % (self.x, self.y)

    def __eq__(self, o):
        try:
            return self.x == o[0] and self.y == o[1] and len(o) == 2
        except Exception:
            return False

    def __ne__(self, o):
        return not self.__eq__(o)

    def __add__(self, o):
        return Vec2d(self.x + o[0], self.y + o[1])
    __radd__ = __add__

    def __sub__(self, o):
        return Vec2d(self.x - o[0], self.y - o[1])

    def __rsub__(self, o):
        return Vec2d(o[0] - self.x, o[1] - self.y)

    def __mul__(self, s):
        return Vec2d(self.x * s, self.y * s)
    __rmul__ = __mul__

    def __truediv__(self, s):
        return Vec2d(self.x / s, self.y / s)

    def __neg__(self):
        return Vec2d(-self.x, -self.y)

    def get_distance(self, o):
        return math.sqrt((self.x - o[0]) ** 2 + (self.y - o[1]) ** 2)

    def get_length(self):
        return math.sqrt(self.x * self.x + self.y * self.y)
    length = property(get_length)

    def dot(self, o):
        return self.x * o[0] + self.y * o[1]

    def cross(self, o):
        return self.x * o[1] - self.y * o[0]

    @property
    def int_tuple(self):
        return int(self.x), int(self.y)


def _dot(ax, ay, bx, by):
    return ax * bx + ay * by


def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


# ------------------------------------------------------------------------------------ BB / Transform
class BB(object):
    def __init__(self, left=0.0, bottom=0.0, right=0.0, top=0.0):
        self.left, self.bottom, self.right, self.top = left, bottom, right, top

    def center(self):
        # cpBBCenter: lerp(lb, rt, 0.5)
        return Vec2d(self.left * 0.5 + self.right * 0.5, self.bottom * 0.5 + self.top * 0.5)

    def intersects(self, o):
        # cpBBIntersects: inclusive
        return self.left <= o.right and o.left <= self.right and self.bottom <= o.top and o.bottom <= self.top

    def merge(self, o):
        return BB(min(self.left, o.left), min(self.bottom, o.bottom), max(self.right, o.right), max(self.top, o.top))

    def __repr__(self):
        return "BB(%r, %r, %r, %r)" % (self.left, self.bottom, self.right, self.top)


class Transform(object):
    def __init__(self, a=1, b=0, c=0, d=1, tx=0, ty=0):
        self.a, self.b, self.c, self.d, self.tx, self.ty = a, b, c, d, tx, ty

    @staticmethod
    def identity():
        return Transform()


class ShapeFilter(object):
    ALL_MASKS = 0xFFFFFFFF
    ALL_CATEGORIES = 0xFFFFFFFF

    def __init__(self, group=0, categories=0xFFFFFFFF, mask=0xFFFFFFFF):
        self.group, self.categories, self.mask = group, categories, mask

    def rejects(self, o):
        # cpShapeFilterReject (chipmunk_private.h)
        return ((self.group != 0 and self.group == o.group)
                or (self.categories & o.mask) == 0
                or (o.categories & self.mask) == 0)


class SpaceDebugDrawOptions(object):
    DRAW_SHAPES = 1
    DRAW_CONSTRAINTS = 2
    DRAW_COLLISION_POINTS = 4


# --------------------------------------------------------------------------------- moments (chipmunk.c)
def moment_for_circle(mass, inner_radius, outer_radius, offset=(0, 0)):
    # cpMomentForCircle: m*(0.5*(r1^2 + r2^2) + |offset|^2)
    return mass * (0.5 * (inner_radius * inner_radius + outer_radius * outer_radius)
                   + (offset[0] * offset[0] + offset[1] * offset[1]))


def moment_for_poly(mass, vertices, offset=(0, 0), radius=0):
    # cpMomentForPoly: m*sum(a_i*b_i)/(6*sum(a_i)), a_i = cross(v2, v1), b_i = v1.v1 + v1.v2 + v2.v2
    vs = [(v[0] + offset[0], v[1] + offset[1]) for v in vertices]
    n = len(vs)
    sum1 = 0.0
    sum2 = 0.0
    for i in range(n):
        v1 = vs[i]
        v2 = vs[(i + 1) % n]
        a = _cross(v2[0], v2[1], v1[0], v1[1])
        b = _dot(*v1, *v1) + _dot(*v1, *v2) + _dot(*v2, *v2)
        sum1 += a * b
        sum2 += a
    return (mass * sum1) / (6.0 * sum2)


# ------------------------------------------------------------------------------- convex hull (cpPolyline/chipmunk.c)
def convex_hull(points):
    """cpConvexHull(count, verts, result, first, tol=0.0): counter-clockwise hull, points that are
    collinear with (or inside) a hull edge are dropped.  Chipmunk uses QuickHull; the vertex SET of a
    tol=0 hull is algorithm independent, and the rotation of the vertex list only changes the order in
    which planes are visited (immaterial away from exact ties).  We start at the min-x (then min-y) vertex
    like cpLoopIndexes does."""
    pts = sorted(set((float(p[0]), float(p[1])) for p in points))
    if len(pts) <= 2:
        return pts

    def half(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and _cross(h[-1][0] - h[-2][0], h[-1][1] - h[-2][1], p[0] - h[-2][0], p[1] - h[-2][1]) <= 0.0:
                h.pop()
            h.append(p)
        return h

    lower = half(pts)
    upper = half(reversed(pts))
    return lower[:-1] + upper[:-1]          # CCW, starts at min-x/min-y


# --------------------------------------------------------------------------------------------- Body
class Body(object):
    DYNAMIC = 0
    KINEMATIC = 1
    STATIC = 2

    def __init__(self, mass=0, moment=0, body_type=DYNAMIC):
        self.body_type = body_type
        if body_type == Body.DYNAMIC:
            self.mass = float(mass)
            self.moment = float(moment)
            self._m_inv = 1.0 / self.mass if self.mass else 0.0
            self._i_inv = 1.0 / self.moment if self.moment else 0.0
        else:
            self.mass = inf
            self.moment = inf
            self._m_inv = 0.0
            self._i_inv = 0.0
        self._p = Vec2d(0.0, 0.0)
        self._v = Vec2d(0.0, 0.0)
        self._f = Vec2d(0.0, 0.0)
        self._cog = Vec2d(0.0, 0.0)
        self.angle = 0.0
        self.angular_velocity = 0.0
        self.torque = 0.0
        self.shapes = set()
        self.space = None

    # pymunk returns a fresh Vec2d from each getter
    def _get_position(self):
        return Vec2d(self._p.x, self._p.y)

    def _set_position(self, p):
        self._p = Vec2d(float(p[0]), float(p[1]))
    position = property(_get_position, _set_position)

    def _get_velocity(self):
        return Vec2d(self._v.x, self._v.y)

    def _set_velocity(self, v):
        self._v = Vec2d(float(v[0]), float(v[1]))
    velocity = property(_get_velocity, _set_velocity)

    def _get_force(self):
        return Vec2d(self._f.x, self._f.y)

    def _set_force(self, f):
        self._f = Vec2d(float(f[0]), float(f[1]))
    force = property(_get_force, _set_force)

    @property
    def center_of_gravity(self):
        return Vec2d(self._cog.x, self._cog.y)

    @property
    def rotation_vector(self):
        return Vec2d(math.cos(self.angle), math.sin(self.angle))

    def _rot(self, v):
        c, s = math.cos(self.angle), math.sin(self.angle)
        return (v[0] * c - v[1] * s, v[0] * s + v[1] * c)

    def local_to_world(self, v):
        r = self._rot(v)
        return Vec2d(r[0] + self._p.x, r[1] + self._p.y)

    def apply_force_at_world_point(self, force, point):
        # cpBodyApplyForceAtWorldPoint: f += force; r = point - transform(cog); t += cross(r, force)
        self._f = Vec2d(self._f.x + force[0], self._f.y + force[1])
        cw = self.local_to_world(self._cog)
        rx, ry = point[0] - cw.x, point[1] - cw.y
        self.torque += _cross(rx, ry, force[0], force[1])

    def apply_force_at_local_point(self, force, point=(0, 0)):
        # cpBodyApplyForceAtLocalPoint: world-ise both arguments, then the world-point version
        fw = self._rot(force)
        pw = self.local_to_world(point)
        self.apply_force_at_world_point(fw, pw)

    # cpBodyUpdatePosition / cpBodyUpdateVelocity (cpBody.c)
    def _update_position(self, dt):
        self._p = Vec2d(self._p.x + self._v.x * dt, self._p.y + self._v.y * dt)
        self.angle = self.angle + self.angular_velocity * dt

    def _update_velocity(self, gravity, damping, dt):
        self._v = Vec2d(self._v.x * damping + (gravity[0] + self._f.x * self._m_inv) * dt,
                        self._v.y * damping + (gravity[1] + self._f.y * self._m_inv) * dt)
        self.angular_velocity = self.angular_velocity * damping + self.torque * self._i_inv * dt
        self._f = Vec2d(0.0, 0.0)
        self.torque = 0.0


# -------------------------------------------------------------------------------------------- Shapes
class SegmentQueryInfo(object):
    def __init__(self, shape, point, normal, alpha):
        self.shape, self.point, self.normal, self.alpha = shape, point, normal, alpha

    def __repr__(self):
        return "SegmentQueryInfo(%r, %r, %r, %r)" % (self.shape, self.point, self.normal, self.alpha)


class PointQueryInfo(object):
    def __init__(self, shape, point, distance, gradient):
        self.shape, self.point, self.distance, self.gradient = shape, point, distance, gradient


def _closest_point_on_segment(px, py, ax, ay, bx, by):
    # cpClosetPointOnSegment (chipmunk_private / cpVect.h)
    dx, dy = ax - bx, ay - by
    den = dx * dx + dy * dy
    t = ((dx * (px - bx) + dy * (py - by)) / den) if den != 0.0 else 0.0
    t = min(max(t, 0.0), 1.0)
    return bx + dx * t, by + dy * t


def _circle_segment_query(shape, cx, cy, r1, a, b, r2, info):
    # CircleSegmentQuery (chipmunk_private.h).  `info` is updated in place on a hit.
    dax, day = a[0] - cx, a[1] - cy
    dbx, dby = b[0] - cx, b[1] - cy
    rsum = r1 + r2
    qa = _dot(dax, day, dax, day) - 2.0 * _dot(dax, day, dbx, dby) + _dot(dbx, dby, dbx, dby)
    qb = _dot(dax, day, dbx, dby) - _dot(dax, day, dax, day)
    det = qb * qb - qa * (_dot(dax, day, dax, day) - rsum * rsum)
    if det >= 0.0 and qa != 0.0:
        t = (-qb - math.sqrt(det)) / qa
        if 0.0 <= t <= 1.0:
            nx, ny = dax + (dbx - dax) * t, day + (dby - day) * t
            ln = math.sqrt(nx * nx + ny * ny)
            if ln > 0.0:
                nx, ny = nx / ln, ny / ln
            info.shape = shape
            info.point = Vec2d(a[0] + (b[0] - a[0]) * t - nx * r2, a[1] + (b[1] - a[1]) * t - ny * r2)
            info.normal = Vec2d(nx, ny)
            info.alpha = t


class Shape(object):
    _next_id = 0

    def __init__(self, body):
        self.body = body
        self.collision_type = 0
        self.filter = ShapeFilter()
        self.friction = 0.0
        self.elasticity = 0.0
        self.sensor = False
        self.color = None
        self.space = None
        self._bb = BB(0.0, 0.0, 0.0, 0.0)   # cpShapeInit leaves bb zeroed (calloc) until cpShapeUpdate
        self._id = Shape._next_id
        Shape._next_id += 1
        if body is not None:
            body.shapes.add(self)

    @property
    def bb(self):
        # cpShapeGetBB: the CACHED box (last cpShapeUpdate / cpShapeCacheBB), not a fresh one
        return BB(self._bb.left, self._bb.bottom, self._bb.right, self._bb.top)

    def cache_bb(self):
        return self.update(None)

    def segment_query(self, start, end, radius=0):
        # cpShapeSegmentQuery (cpShape.c) + pymunk Shape.segment_query wrapping
        a = (float(start[0]), float(start[1]))
        b = (float(end[0]), float(end[1]))
        info = SegmentQueryInfo(None, Vec2d(b[0], b[1]), Vec2d(0.0, 0.0), 1.0)
        nearest = self.point_query(a)
        if nearest.distance <= radius:
            info.shape = self
            info.alpha = 0.0
            nx, ny = a[0] - nearest.point.x, a[1] - nearest.point.y
            ln = math.sqrt(nx * nx + ny * ny)
            info.normal = Vec2d(nx / ln, ny / ln) if ln > 0.0 else Vec2d(0.0, 0.0)
        else:
            self._segment_query(a, b, radius, info)
        return info


class Circle(Shape):
    def __init__(self, body, radius, offset=(0, 0)):
        Shape.__init__(self, body)
        self.radius = float(radius)
        self.offset = Vec2d(float(offset[0]), float(offset[1]))
        self._tc = Vec2d(0.0, 0.0)

    def update(self, transform):
        # cpCircleShapeCacheData
        c = self.body.local_to_world(self.offset)
        self._tc = c
        r = self.radius
        self._bb = BB(c.x - r, c.y - r, c.x + r, c.y + r)
        return self.bb

    def point_query(self, p):
        # cpCircleShapePointQuery
        dx, dy = p[0] - self._tc.x, p[1] - self._tc.y
        d = math.sqrt(dx * dx + dy * dy)
        r = self.radius
        if d > 0.0:
            pt = Vec2d(self._tc.x + dx * (r / d), self._tc.y + dy * (r / d))
            g = Vec2d(dx / d, dy / d)
        else:
            pt = Vec2d(self._tc.x, self._tc.y)
            g = Vec2d(0.0, 1.0)
        return PointQueryInfo(self, pt, d - r, g)

    def _segment_query(self, a, b, radius, info):
        _circle_segment_query(self, self._tc.x, self._tc.y, self.radius, a, b, radius, info)


class Poly(Shape):
    def __init__(self, body, vertices, transform=None, radius=0):
        Shape.__init__(self, body)
        # cpPolyShapeInit: convex hull (tol 0) of the given vertices, CCW; SetVerts builds the planes
        self._local = convex_hull(vertices)
        self.radius = float(radius)
        n = len(self._local)
        self._count = n
        self._verts = [(0.0, 0.0)] * n      # world vertices  (planes[i].v0)
        self._normals = [(0.0, 0.0)] * n    # world normals   (planes[i].n): plane i is edge v[i-1] -> v[i]

    def get_vertices(self):
        return [Vec2d(v) for v in self._local]

    def update(self, transform):
        # cpPolyShapeCacheData: transform hull verts and normals, AABB = min/max of verts (+radius)
        body = self.body
        c, s = math.cos(body.angle), math.sin(body.angle)
        px, py = body._p.x, body._p.y
        w = []
        for (x, y) in self._local:
            w.append((x * c - y * s + px, x * s + y * c + py))
        n = self._count
        normals = []
        for i in range(n):
            ax, ay = w[(i - 1 + n) % n]
            bx, by = w[i]
            ex, ey = bx - ax, by - ay
            # cpvrperp(e) = (e.y, -e.x): outward for a CCW loop.  (Chipmunk rotates the precomputed
            # local normal instead; identical up to rounding.)
            ln = math.sqrt(ex * ex + ey * ey)
            normals.append((ey / ln, -ex / ln))
        self._verts = w
        self._normals = normals
        r = self.radius
        self._bb = BB(min(v[0] for v in w) - r, min(v[1] for v in w) - r,
                      max(v[0] for v in w) + r, max(v[1] for v in w) + r)
        return self.bb

    def point_query(self, p):
        # cpPolyShapePointQuery (cpPolyShape.c)
        n = self._count
        w, normals, r = self._verts, self._normals, self.radius
        v0 = w[n - 1]
        min_dist = inf
        closest = (0.0, 0.0)
        closest_n = (0.0, 0.0)
        outside = False
        for i in range(n):
            v1 = w[i]
            outside = outside or (_dot(normals[i][0], normals[i][1], p[0] - v1[0], p[1] - v1[1]) > 0.0)
            cx, cy = _closest_point_on_segment(p[0], p[1], v0[0], v0[1], v1[0], v1[1])
            d = math.sqrt((p[0] - cx) ** 2 + (p[1] - cy) ** 2)
            if d < min_dist:
                min_dist = d
                closest = (cx, cy)
                closest_n = normals[i]
            v0 = v1
        dist = min_dist if outside else -min_dist
        if dist != 0.0:
            g = ((p[0] - closest[0]) / dist, (p[1] - closest[1]) / dist)
        else:
            g = closest_n
        return PointQueryInfo(self, Vec2d(closest[0] + g[0] * r, closest[1] + g[1] * r), dist - r, Vec2d(g))

    def _segment_query(self, a, b, r2, info):
        # cpPolyShapeSegmentQuery (cpPolyShape.c)
        n = self._count
        w, normals = self._verts, self._normals
        r = self.radius
        rsum = r + r2
        for i in range(n):
            nx, ny = normals[i]
            an = _dot(a[0], a[1], nx, ny)
            d = an - _dot(w[i][0], w[i][1], nx, ny) - rsum
            if d < 0.0:
                continue
            bn = _dot(b[0], b[1], nx, ny)
            t = d / max(an - bn, DBL_MIN)
            if t < 0.0 or 1.0 < t:
                continue
            px, py = a[0] + (b[0] - a[0]) * t, a[1] + (b[1] - a[1]) * t
            dt = _cross(nx, ny, px, py)
            dt_min = _cross(nx, ny, *w[(i - 1 + n) % n])
            dt_max = _cross(nx, ny, *w[i])
            if dt_min <= dt <= dt_max:
                info.shape = self
                info.point = Vec2d(px - nx * r2, py - ny * r2)
                info.normal = Vec2d(nx, ny)
                info.alpha = t
        if rsum > 0.0:
            for i in range(n):
                ci = SegmentQueryInfo(None, Vec2d(b[0], b[1]), Vec2d(0.0, 0.0), 1.0)
                _circle_segment_query(self, w[i][0], w[i][1], r, a, b, r2, ci)
                if ci.alpha < info.alpha:
                    info.shape, info.point, info.normal, info.alpha = ci.shape, ci.point, ci.normal, ci.alpha


# --------------------------------------------------------------------- narrow phase (cpCollision.c predicates)
def _poly_poly_distance(p1, p2):
    """Signed separation of two convex polygons the way GJK/EPA reports it for the contact test
    `points.d - r1 - r2 <= 0` (PolyToPoly, cpCollision.c): > 0 = Euclidean gap when disjoint,
    <= 0 when touching / penetrating (value = minus the minimum translation distance)."""
    w1, w2 = p1._verts, p2._verts

    def max_sep(wa, na, wb):
        best = -inf
        for i in range(len(wa)):
            nx, ny = na[i]
            off = _dot(wa[i][0], wa[i][1], nx, ny)
            m = min(_dot(v[0], v[1], nx, ny) for v in wb) - off
            if m > best:
                best = m
        return best

    sat = max(max_sep(w1, p1._normals, w2), max_sep(w2, p2._normals, w1))
    if sat <= 0.0:
        return sat
    # disjoint: true distance = min over vertex/edge pairs
    best = inf
    for (wa, wb) in ((w1, w2), (w2, w1)):
        nb = len(wb)
        for (px, py) in wa:
            for j in range(nb):
                ax, ay = wb[j - 1]
                bx, by = wb[j]
                cx, cy = _closest_point_on_segment(px, py, ax, ay, bx, by)
                d = math.sqrt((px - cx) ** 2 + (py - cy) ** 2)
                if d < best:
                    best = d
    return best


def _shapes_touch(a, b):
    """info.count > 0 of cpCollide(a, b): CircleToPoly `d <= r_c + r_p`, PolyToPoly `d - r1 - r2 <= 0`,
    CircleToCircle `d < r1 + r2` (strict, see CircleToCircleQuery)."""
    if isinstance(a, Circle) and isinstance(b, Circle):
        return a._tc.get_distance(b._tc) < a.radius + b.radius
    if isinstance(a, Circle) and isinstance(b, Poly):
        return b.point_query((a._tc.x, a._tc.y)).distance <= a.radius
    if isinstance(a, Poly) and isinstance(b, Circle):
        return a.point_query((b._tc.x, b._tc.y)).distance <= b.radius
    return _poly_poly_distance(a, b) - a.radius - b.radius <= 0.0


class Arbiter(object):
    def __init__(self, a, b):
        self.shapes = (a, b)
        self.is_first_contact = True


class CollisionHandler(object):
    def __init__(self, type_a, type_b):
        self.type_a, self.type_b = type_a, type_b
        self.begin = None
        self.pre_solve = None
        self.post_solve = None
        self.separate = None
        self.data = {}


# --------------------------------------------------------------------------------------------- Space
class Space(object):
    def __init__(self, threaded=False):
        self.damping = 1.0
        self.gravity = Vec2d(0.0, 0.0)
        self.iterations = 10
        self.static_body = Body(body_type=Body.STATIC)
        self._static_shapes = []        # insertion order, queried first (cpSpaceSegmentQuery)
        self._dynamic_shapes = []
        self._bodies = []
        self._handlers = {}
        self._cached_pairs = {}          # (id_a, id_b) -> accepted? ; persists while the pair keeps touching
        self._locked = False
        self._deferred_remove = []
        self.current_time_step = 0.0

    @property
    def shapes(self):
        return list(self._static_shapes) + list(self._dynamic_shapes)

    @property
    def bodies(self):
        return list(self._bodies)

    def add(self, *objs):
        for o in objs:
            if isinstance(o, Body):
                o.space = self
                if o.body_type == Body.DYNAMIC:
                    self._bodies.append(o)
            elif isinstance(o, Shape):
                o.space = self
                o.update(None)           # cpSpaceAddShape -> cpShapeUpdate(shape, body->transform)
                if o.body.body_type == Body.STATIC:
                    self._static_shapes.append(o)
                else:
                    self._dynamic_shapes.append(o)
            else:
                raise TypeError(o)

    def remove(self, *objs):
        # pymunk defers add/remove while the space is locked (inside step callbacks)
        if self._locked:
            self._deferred_remove.extend(objs)
            return
        for o in objs:
            if isinstance(o, Body):
                if o in self._bodies:
                    self._bodies.remove(o)
                o.space = None
            elif isinstance(o, Shape):
                if o in self._static_shapes:
                    self._static_shapes.remove(o)
                if o in self._dynamic_shapes:
                    self._dynamic_shapes.remove(o)
                for k in [k for k in self._cached_pairs if o._id in k]:
                    del self._cached_pairs[k]
                o.space = None

    def add_collision_handler(self, a, b):
        key = (min(a, b), max(a, b))
        if key not in self._handlers:
            self._handlers[key] = CollisionHandler(a, b)
        return self._handlers[key]

    def debug_draw(self, options):
        pass

    def segment_query(self, start, end, radius, shape_filter):
        # cpSpaceSegmentQuery: static index first, then dynamic; every shape that reports a hit is
        # appended (pymunk 5.4 returns the unsorted list).  The spatial-index pre-filter tests the THIN
        # segment against each leaf's BB (cpBBTree SubtreeSegmentQuery -> cpBBSegmentQuery).
        a = (float(start[0]), float(start[1]))
        b = (float(end[0]), float(end[1]))
        hits = []
        for s in list(self._static_shapes) + list(self._dynamic_shapes):
            if s.filter.rejects(shape_filter):
                continue
            if not _bb_segment_hits(s._bb, a, b):
                continue
            info = s.segment_query(a, b, radius)
            if info.shape is not None:
                hits.append(info)
        return hits

    def step(self, dt):
        # cpSpaceStep (cpSpaceStep.c), in Chipmunk's order
        if dt == 0.0:
            return
        self.current_time_step = dt
        self._locked = True
        # (1) integrate positions
        for b in self._bodies:
            b._update_position(dt)
        # (2) refresh cached world verts / planes / BBs of dynamic shapes
        for s in self._dynamic_shapes:
            s.update(None)
        # (3) find colliding pairs; `begin` fires on the first step a pair touches
        touching = {}
        dyn = list(self._dynamic_shapes)
        others = list(self._static_shapes)
        for i, a in enumerate(dyn):
            for b in others + dyn[i + 1:]:
                if a.body is b.body or not a._bb.intersects(b._bb) or a.filter.rejects(b.filter):
                    continue
                if not _shapes_touch(a, b):
                    continue
                key = (min(a._id, b._id), max(a._id, b._id))
                if key in self._cached_pairs:
                    touching[key] = self._cached_pairs[key]
                    continue
                h = self._handlers.get((min(a.collision_type, b.collision_type),
                                        max(a.collision_type, b.collision_type)))
                accepted = True
                if h is not None and h.begin is not None:
                    # order arbiter.shapes like the handler's (type_a, type_b) -- cpArbiterUpdate `swapped`
                    sa, sb = (a, b) if a.collision_type == h.type_a else (b, a)
                    accepted = bool(h.begin(Arbiter(sa, sb), self, h.data))
                touching[key] = accepted
        self._cached_pairs = touching      # pairs that separated are forgotten -> `begin` can fire again
        # (4) integrate velocities: damping = pow(space.damping, dt)
        damping = math.pow(self.damping, dt)
        for b in self._bodies:
            b._update_velocity(self.gravity, damping, dt)
        # (5) impulse solver: NOT restated (see module docstring)
        self._locked = False
        # (6) deferred removals
        if self._deferred_remove:
            objs, self._deferred_remove = self._deferred_remove, []
            self.remove(*objs)


def _bb_segment_hits(bb, a, b):
    # cpBBSegmentQuery(bb, a, b) != INFINITY  (slab test, thin segment)
    dx, dy = b[0] - a[0], b[1] - a[1]
    tmin, tmax = -inf, inf
    if dx == 0.0:
        if a[0] < bb.left or bb.right < a[0]:
            return False
    else:
        t1, t2 = (bb.left - a[0]) / dx, (bb.right - a[0]) / dx
        tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
    if dy == 0.0:
        if a[1] < bb.bottom or bb.top < a[1]:
            return False
    else:
        t1, t2 = (bb.bottom - a[1]) / dy, (bb.top - a[1]) / dy
        tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
    return tmin <= tmax and 0.0 <= tmax and tmin <= 1.0
