/*
 * shipsim_oracle.c -- float64 CPU restatement of the CapAI/ship-sim-gym ShipEnv transition.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (ship_sim_gym_b200/) never links, imports or calls it and has no CPU fallback.
 *
 * Pinning status.  The reference's own Python logic is pinned: the tests/golden fixtures were produced by
 * executing the unmodified reference (ship_gym modules imported from /root/reference) over a pure-Python
 * restatement of the pymunk/Chipmunk calls it makes (oracle/shims/, oracle/make_golden.py), and this C
 * file reproduces those fixtures to 1e-9 (tests/test_oracle_golden.py).  The Chipmunk2D 7.0.2
 * arithmetic underneath (pymunk==5.4.0, requirements.txt:78; not vendored, not installable here) is
 * restated from its published algorithm and has NOT been compared with a live libchipmunk:
 * "Chipmunk layer: parity unpinned".
 *
 * Reference map (file:line under /root/reference) is given at each function.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_MAXV 40        /* max hull vertices per bank */
#define ORC_NGOALS 5       /* game.py:17 N_GOALS */
#define ORC_MAXBEAMS 32
#define ORC_FRAME(nb) (6 + (nb))   /* ship_env.py:43  n_states = 2+1+1+2+n_beams */
#define ORC_MAXHIST 8

typedef struct {
    double W, H;                 /* config.py:24 BOUNDS */
    double dt;                   /* game.py:194 speed*base_dt */
    double space_damping;        /* game.py:270 (0.4); per-step factor is pow(space_damping, dt) */
    double lidar_spread_deg;     /* models.py:29 (90) */
    double lidar_distance;       /* models.py:29 (100) */
    double ship_w, ship_h;       /* game.py:275 add_player_ship(..., 2, 3, ...) */
    double mass;                 /* models.py:87 (5) */
    double thrust;               /* models.py:107 force_vector=(0,100) */
    double goal_radius;          /* game.py:82 (5) */
    double step_penalty;         /* ship_env.py:13 (-0.01) */
    double spawn_y;              /* game.py:274 (25) */
    uint64_t seed;               /* scenario choice on auto-reset: hash(seed, global env id, episode) */
    int64_t env_id_offset;       /* global id of env 0 (multi-GPU sharding) */
    int32_t max_steps;           /* config.py:16 */
    int32_t history;             /* config.py:15 */
    int32_t n_beams;             /* models.py:29 (10) */
    int32_t auto_reset;          /* SubprocVecEnv worker semantics */
    int32_t n_scenarios;
    int32_t maxv;                /* row stride of the hull arrays (<= ORC_MAXV) */
    int32_t pick_base;           /* fresh-maps mode of the product (pick_count > 0, a power of two): resets walk the bank slice */
    int32_t pick_count;          /* [pick_base, pick_base + pick_count) with a per-env offset and odd stride                  */
} orc_config;

/* scenario bank, flat: hull_xy[s][b][maxv][2], hull_n[s][b], goals[s][5][2] */
typedef struct {
    const double *hull_xy;
    const int32_t *hull_n;
    const double *goals;
} orc_bank;

/* per-env state, SoA of small rows (double / int32) */
typedef struct {
    double *pose;      /* [N][6]  x y angle vx vy w                         */
    double *lidar;     /* [N][ORC_MAXBEAMS] sticky LiDAR.vals (models.py:36) */
    double *goals;     /* [N][5][2]                                         */
    double *ep_return; /* [N]    ship_env.py:150 cumulative_reward          */
    double *hist;      /* [N][history*frame] the deque of ship_env.py:181   */
    int32_t *ints;     /* [N][5]  rudder, alive_mask, step_count, scenario, episode */
} orc_state;

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) -- the counter RNG shared with the CUDA kernel.
 * ---------------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32(uint64_t seed, uint64_t ctr_lo, uint32_t ctr2, uint32_t ctr3, uint32_t out[4])
{
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), ctr2, ctr3};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}

/* stream ids (counter word 3) */
#define ORC_STREAM_SCENARIO 0u
#define ORC_STREAM_ACTION 1u

/* Scenario picks: the product's bookkeeping (the reference builds a new level per reset, game.py:271-272, and has no
 * bank), restated so that auto-reset trajectories can be compared: an integer hash of (seed, global env id, episode). */
static uint32_t mix32(uint32_t x)
{
    x ^= x >> 17; x *= 0xed5ad4bbu; x ^= x >> 11; x *= 0xac4c1b51u; x ^= x >> 15; x *= 0x31848babu; x ^= x >> 14;
    return x;
}

static uint32_t pick_key(uint64_t seed, int64_t gid)
{
    uint32_t k = mix32((uint32_t)seed ^ 0x9E3779B9u);
    k = mix32(k ^ (uint32_t)(seed >> 32));
    k = mix32(k ^ (uint32_t)(uint64_t)gid);
    return mix32(k ^ (uint32_t)((uint64_t)gid >> 32));
}

int32_t orc_pick_scenario(uint64_t seed, int64_t gid, int32_t episode, int32_t n_scenarios)
{
    const uint32_t h = mix32(pick_key(seed, gid) + (uint32_t)episode * 0x9E3779B9u);
    return (int32_t)(((uint64_t)h * (uint64_t)n_scenarios) >> 32);
}

/* fresh-maps mode: a walk through the newest slice of the bank with a per-env offset and odd stride */
static int32_t pick_for(const orc_config *c, int64_t gid, int32_t episode)
{
    if (c->pick_count > 0) {
        const uint32_t key = pick_key(c->seed, gid);
        const uint32_t stride = mix32(key ^ 0x5bd1e995u) | 1u;
        return c->pick_base + (int32_t)((key + (uint32_t)episode * stride) & (uint32_t)(c->pick_count - 1));
    }
    return orc_pick_scenario(c->seed, gid, episode, c->n_scenarios);
}

int32_t orc_random_action(uint64_t seed, int64_t gid, uint32_t step)
{
    /* uniform{0,1,2}: the random agent of train/random.py:18 (env.action_space.sample()) */
    uint32_t r[4];
    orc_philox4x32(seed, (uint64_t)gid, step, ORC_STREAM_ACTION, r);
    return (int32_t)(((uint64_t)r[0] * 3u) >> 32);
}

/* ------------------------------------------------------------------------------------------------
 * Chipmunk2D 7.0.2 pieces (restated; see header)
 * ---------------------------------------------------------------------------------------------- */

/* cpMomentForPoly (chipmunk.c); call site models.py:89 */
double orc_moment_for_poly(double m, int n, const double *xy)
{
    double sum1 = 0.0, sum2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x1 = xy[2 * i], y1 = xy[2 * i + 1];
        const double x2 = xy[2 * ((i + 1) % n)], y2 = xy[2 * ((i + 1) % n) + 1];
        const double a = x2 * y1 - y2 * x1;
        const double b = (x1 * x1 + y1 * y1) + (x1 * x2 + y1 * y2) + (x2 * x2 + y2 * y2);
        sum1 += a * b;
        sum2 += a;
    }
    return (m * sum1) / (6.0 * sum2);
}

static int cmp_pt(const void *a, const void *b)
{
    const double *p = (const double *)a, *q = (const double *)b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    if (p[1] != q[1]) return p[1] < q[1] ? -1 : 1;
    return 0;
}

/* cpConvexHull(count, verts, result, NULL, tol=0): CCW hull, collinear points dropped; what pm.Poly does
 * to its vertex list (models.py:96,180).  Starts at the min-x (then min-y) vertex. */
int orc_convex_hull(int n, const double *xy_in, double *xy_out)
{
    double pts[2 * 64];
    double h[2 * 130];
    if (n > 64) return -1;
    memcpy(pts, xy_in, sizeof(double) * 2 * (size_t)n);
    qsort(pts, (size_t)n, 2 * sizeof(double), cmp_pt);
    int m = 0;
    for (int i = 0; i < n; ++i)          /* drop exact duplicates */
        if (m == 0 || pts[2 * i] != pts[2 * (m - 1)] || pts[2 * i + 1] != pts[2 * (m - 1) + 1]) {
            pts[2 * m] = pts[2 * i]; pts[2 * m + 1] = pts[2 * i + 1]; ++m;
        }
    n = m;
    if (n <= 2) { memcpy(xy_out, pts, sizeof(double) * 2 * (size_t)n); return n; }
    int k = 0;
    for (int i = 0; i < n; ++i) {        /* lower chain */
        while (k >= 2) {
            const double ax = h[2 * (k - 1)] - h[2 * (k - 2)], ay = h[2 * (k - 1) + 1] - h[2 * (k - 2) + 1];
            const double bx = pts[2 * i] - h[2 * (k - 2)], by = pts[2 * i + 1] - h[2 * (k - 2) + 1];
            if (ax * by - ay * bx <= 0.0) --k; else break;
        }
        h[2 * k] = pts[2 * i]; h[2 * k + 1] = pts[2 * i + 1]; ++k;
    }
    const int lower = k + 1;
    for (int i = n - 2; i >= 0; --i) {   /* upper chain */
        while (k >= lower) {
            const double ax = h[2 * (k - 1)] - h[2 * (k - 2)], ay = h[2 * (k - 1) + 1] - h[2 * (k - 2) + 1];
            const double bx = pts[2 * i] - h[2 * (k - 2)], by = pts[2 * i + 1] - h[2 * (k - 2) + 1];
            if (ax * by - ay * bx <= 0.0) --k; else break;
        }
        h[2 * k] = pts[2 * i]; h[2 * k + 1] = pts[2 * i + 1]; ++k;
    }
    --k;                                  /* last point repeats the first */
    memcpy(xy_out, h, sizeof(double) * 2 * (size_t)k);
    return k;
}

typedef struct {
    int n;
    double vx[ORC_MAXV], vy[ORC_MAXV];   /* planes[i].v0 */
    double nx[ORC_MAXV], ny[ORC_MAXV];   /* planes[i].n : outward normal of edge v[i-1] -> v[i] */
} orc_poly;

/* SetVerts / cpPolyShapeCacheData (cpPolyShape.c): planes from a CCW loop */
static void poly_from_world(orc_poly *p, int n, const double *xy)
{
    p->n = n;
    for (int i = 0; i < n; ++i) { p->vx[i] = xy[2 * i]; p->vy[i] = xy[2 * i + 1]; }
    for (int i = 0; i < n; ++i) {
        const int j = (i - 1 + n) % n;
        const double ex = p->vx[i] - p->vx[j], ey = p->vy[i] - p->vy[j];
        const double ln = sqrt(ex * ex + ey * ey);
        p->nx[i] = ey / ln;              /* cpvrperp */
        p->ny[i] = -ex / ln;
    }
}

static void closest_on_segment(double px, double py, double ax, double ay, double bx, double by,
                               double *cx, double *cy)
{
    /* cpClosetPointOnSegment (cpVect.h) */
    const double dx = ax - bx, dy = ay - by;
    const double den = dx * dx + dy * dy;
    double t = den != 0.0 ? (dx * (px - bx) + dy * (py - by)) / den : 0.0;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    *cx = bx + dx * t;
    *cy = by + dy * t;
}

/* cpPolyShapePointQuery (cpPolyShape.c): signed distance, negative inside */
static double poly_point_distance(const orc_poly *p, double px, double py)
{
    double min_dist = INFINITY;
    int outside = 0;
    double v0x = p->vx[p->n - 1], v0y = p->vy[p->n - 1];
    for (int i = 0; i < p->n; ++i) {
        const double v1x = p->vx[i], v1y = p->vy[i];
        if (p->nx[i] * (px - v1x) + p->ny[i] * (py - v1y) > 0.0) outside = 1;
        double cx, cy;
        closest_on_segment(px, py, v0x, v0y, v1x, v1y, &cx, &cy);
        const double d = sqrt((px - cx) * (px - cx) + (py - cy) * (py - cy));
        if (d < min_dist) min_dist = d;
        v0x = v1x; v0y = v1y;
    }
    return outside ? min_dist : -min_dist;
}

typedef struct { int hit; double px, py, alpha, margin, cond; } orc_seg_info;   /* cond = 1/|n.dir| of the accepted edge */

/* CircleSegmentQuery (chipmunk_private.h) */
static void circle_segment_query(double cx, double cy, double r1, double ax, double ay, double bx, double by,
                                 double r2, orc_seg_info *info)
{
    const double dax = ax - cx, day = ay - cy, dbx = bx - cx, dby = by - cy;
    const double rsum = r1 + r2;
    const double daa = dax * dax + day * day, dab = dax * dbx + day * dby, dbb = dbx * dbx + dby * dby;
    const double qa = daa - 2.0 * dab + dbb;
    const double qb = dab - daa;
    const double det = qb * qb - qa * (daa - rsum * rsum);
    if (det >= 0.0 && qa != 0.0) {
        const double t = (-qb - sqrt(det)) / qa;
        if (0.0 <= t && t <= 1.0) {
            double nx = dax + (dbx - dax) * t, ny = day + (dby - day) * t;
            const double ln = sqrt(nx * nx + ny * ny);
            if (ln > 0.0) { nx /= ln; ny /= ln; }
            info->hit = 1;
            info->px = ax + (bx - ax) * t - nx * r2;
            info->py = ay + (by - ay) * t - ny * r2;
            info->alpha = t;
        }
    }
}

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* cpShapeSegmentQuery (cpShape.c) + cpPolyShapeSegmentQuery (cpPolyShape.c); call sites models.py:67
 * (r = 0) and game.py:322-323 (r = 10).  `margin` = smallest slack (in length units) of any inequality
 * that was evaluated on the way to the answer: the grazing filter of SURVEY.md §8(d). */
static void poly_segment_query(const orc_poly *p, double ax, double ay, double bx, double by, double r2,
                               orc_seg_info *info)
{
    info->hit = 0; info->px = bx; info->py = by; info->alpha = 1.0; info->margin = INFINITY; info->cond = 1.0;
    const double nearest = poly_point_distance(p, ax, ay);
    info->margin = dmin(info->margin, fabs(nearest - r2));
    if (nearest <= r2) {               /* start inside (or within r): hit, alpha 0, point stays at b */
        info->hit = 1;
        info->alpha = 0.0;
        return;
    }
    const double len = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
    const int n = p->n;
    for (int i = 0; i < n; ++i) {
        const double nx = p->nx[i], ny = p->ny[i];
        const double an = ax * nx + ay * ny;
        const double d = an - (p->vx[i] * nx + p->vy[i] * ny) - r2;
        info->margin = dmin(info->margin, fabs(d));
        if (d < 0.0) continue;
        const double bn = bx * nx + by * ny;
        const double t = d / dmax(an - bn, DBL_MIN);
        if (t <= 2.0) info->margin = dmin(info->margin, dmin(fabs(t), fabs(1.0 - t)) * len);
        if (t < 0.0 || 1.0 < t) continue;
        const double qx = ax + (bx - ax) * t, qy = ay + (by - ay) * t;
        const int j = (i - 1 + n) % n;
        const double dtv = nx * qy - ny * qx;
        const double dt_min = nx * p->vy[j] - ny * p->vx[j];
        const double dt_max = nx * p->vy[i] - ny * p->vx[i];
        info->margin = dmin(info->margin, dmin(fabs(dtv - dt_min), fabs(dtv - dt_max)));
        if (dt_min <= dtv && dtv <= dt_max) {
            info->hit = 1;
            info->px = qx - nx * r2;
            info->py = qy - ny * r2;
            info->alpha = t;
            info->cond = len / dmax(an - bn, DBL_MIN);
        }
    }
    if (r2 > 0.0) {                    /* bevelled vertices */
        for (int i = 0; i < n; ++i) {
            orc_seg_info ci = {0, bx, by, 1.0, INFINITY, 1.0};
            circle_segment_query(p->vx[i], p->vy[i], 0.0, ax, ay, bx, by, r2, &ci);
            if (ci.alpha < info->alpha) { info->hit = ci.hit; info->px = ci.px; info->py = ci.py; info->alpha = ci.alpha; }
        }
    }
}

/* Narrow phase PolyToPoly (cpCollision.c): contact iff GJK distance <= 0.  Evaluated WITHOUT a
 * separating-axis shortcut so that it is independent of the kernel's SAT: two convex polygons are in
 * contact iff some pair of edges intersects/touches or one contains a vertex of the other.
 * Returns 1/0; *sep receives the SAT separation (>0 gap lower bound, <=0 penetration) as the margin. */
static int seg_seg_touch(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy)
{
    const double d1 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    const double d2 = (bx - ax) * (dy - ay) - (by - ay) * (dx - ax);
    const double d3 = (dx - cx) * (ay - cy) - (dy - cy) * (ax - cx);
    const double d4 = (dx - cx) * (by - cy) - (dy - cy) * (bx - cx);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) return 1;
    /* touching / collinear cases */
    #define ON_SEG(px, py, qx, qy, rx, ry) (dmin(px, qx) <= rx && rx <= dmax(px, qx) && dmin(py, qy) <= ry && ry <= dmax(py, qy))
    if (d1 == 0 && ON_SEG(ax, ay, bx, by, cx, cy)) return 1;
    if (d2 == 0 && ON_SEG(ax, ay, bx, by, dx, dy)) return 1;
    if (d3 == 0 && ON_SEG(cx, cy, dx, dy, ax, ay)) return 1;
    if (d4 == 0 && ON_SEG(cx, cy, dx, dy, bx, by)) return 1;
    #undef ON_SEG
    return 0;
}

static double sat_one_way(const orc_poly *a, const orc_poly *b)
{
    double best = -INFINITY;
    for (int i = 0; i < a->n; ++i) {
        const double off = a->vx[i] * a->nx[i] + a->vy[i] * a->ny[i];
        double m = INFINITY;
        for (int k = 0; k < b->n; ++k) m = dmin(m, b->vx[k] * a->nx[i] + b->vy[k] * a->ny[i] - off);
        best = dmax(best, m);
    }
    return best;
}

static int polys_touch(const orc_poly *a, const orc_poly *b, double *sep)
{
    *sep = dmax(sat_one_way(a, b), sat_one_way(b, a));
    for (int i = 0; i < a->n; ++i) {
        const int i0 = (i - 1 + a->n) % a->n;
        for (int k = 0; k < b->n; ++k) {
            const int k0 = (k - 1 + b->n) % b->n;
            if (seg_seg_touch(a->vx[i0], a->vy[i0], a->vx[i], a->vy[i], b->vx[k0], b->vy[k0], b->vx[k], b->vy[k]))
                return 1;
        }
    }
    if (poly_point_distance(b, a->vx[0], a->vy[0]) <= 0.0) return 1;   /* a inside b */
    if (poly_point_distance(a, b->vx[0], b->vy[0]) <= 0.0) return 1;   /* b inside a */
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * exported geometry entry points (used by tests to cross-check the product's host-side scenario.py)
 * ---------------------------------------------------------------------------------------------- */
double orc_poly_point_distance(int n, const double *hull_xy, double px, double py)
{
    orc_poly p; poly_from_world(&p, n, hull_xy);
    return poly_point_distance(&p, px, py);
}

/* out[5] = hit, point.x, point.y, alpha, margin */
void orc_segment_query(int n, const double *hull_xy, double ax, double ay, double bx, double by, double r,
                       double *out)
{
    orc_poly p; poly_from_world(&p, n, hull_xy);
    orc_seg_info info;
    poly_segment_query(&p, ax, ay, bx, by, r, &info);
    out[0] = info.hit; out[1] = info.px; out[2] = info.py; out[3] = info.alpha; out[4] = info.margin;
}

int orc_polys_touch(int n1, const double *xy1, int n2, const double *xy2, double *sep)
{
    orc_poly a, b; poly_from_world(&a, n1, xy1); poly_from_world(&b, n2, xy2);
    return polys_touch(&a, &b, sep);
}

/* ShipGame.gen_goal_path (game.py:300-330) for ONE goal row: the two fat (r=10) horizontal space
 * segment queries from x=W/2 (game.py:322-323) against the static bank shapes (spatial-index pre-filter:
 * thin segment vs shape BB).  out[2] = left.point.x + 60, right.point.x - 60.  Returns 0 when either
 * query list is empty (the reference's IndexError -> fallback path, game.py:328-330). */
int orc_goal_span(const double *hull_xy, const int32_t *hull_n, int maxv, double W, double y, double *out)
{
    const double tol = 60.0, r = 10.0;
    double res[2];
    for (int side = 0; side < 2; ++side) {
        const double ax = W / 2.0, bx = side == 0 ? 0.0 : W;
        int found = 0;
        for (int b = 0; b < 2 && !found; ++b) {
            const double *xy = hull_xy + (size_t)b * maxv * 2;
            const int n = hull_n[b];
            double l = INFINITY, rt = -INFINITY, bo = INFINITY, tp = -INFINITY;
            for (int i = 0; i < n; ++i) {
                l = dmin(l, xy[2 * i]); rt = dmax(rt, xy[2 * i]);
                bo = dmin(bo, xy[2 * i + 1]); tp = dmax(tp, xy[2 * i + 1]);
            }
            if (y < bo || y > tp) continue;                          /* cpBBSegmentQuery, horizontal ray */
            if (dmax(ax, bx) < l || dmin(ax, bx) > rt) continue;
            orc_poly p; poly_from_world(&p, n, xy);
            orc_seg_info info;
            poly_segment_query(&p, ax, y, bx, y, r, &info);
            if (info.hit) { res[side] = info.px; found = 1; }
        }
        if (!found) return 0;
    }
    out[0] = res[0] + tol;
    out[1] = res[1] - tol;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * The environment transition, one env
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double local[5][2];    /* ship hull in body frame, CCW (convex hull of SHIP_TEMPLATE scaled; models.py:6,88) */
    double moment;         /* models.py:89 */
    double damping;        /* pow(space_damping, dt): cpSpaceStep */
    double ray_cos[ORC_MAXBEAMS], ray_sin[ORC_MAXBEAMS];
} orc_derived;

static void derive(const orc_config *c, orc_derived *d)
{
    /* SHIP_TEMPLATE = [(0,0),(0,10),(5,15),(10,10),(10,0)] (models.py:6) scaled by (width,height) */
    const double tpl[5][2] = {{0, 0}, {0, 10}, {5, 15}, {10, 10}, {10, 0}};
    double pts[10], hull[10];
    for (int i = 0; i < 5; ++i) { pts[2 * i] = tpl[i][0] * c->ship_w; pts[2 * i + 1] = tpl[i][1] * c->ship_h; }
    d->moment = orc_moment_for_poly(c->mass, 5, pts);
    const int n = orc_convex_hull(5, pts, hull);
    (void)n;
    for (int i = 0; i < 5; ++i) { d->local[i][0] = hull[2 * i]; d->local[i][1] = hull[2 * i + 1]; }
    d->damping = pow(c->space_damping, c->dt);
    /* models.py:48-49,62: delta = radians(spread/n), start = angle + radians(90 - spread/2) */
    const double delta = (c->lidar_spread_deg / c->n_beams) * (M_PI / 180.0);
    const double start = (90.0 - c->lidar_spread_deg / 2.0) * (M_PI / 180.0);
    for (int i = 0; i < c->n_beams; ++i) { d->ray_cos[i] = cos(start + delta * i); d->ray_sin[i] = sin(start + delta * i); }
}

static void load_bank_poly(const orc_config *c, const orc_bank *bank, int scen, int b, orc_poly *p)
{
    const double *xy = bank->hull_xy + ((size_t)scen * 2 + b) * c->maxv * 2;
    poly_from_world(p, bank->hull_n[scen * 2 + b], xy);
}

static void ship_world(const orc_derived *d, double x, double y, double th, orc_poly *p)
{
    double xy[10];
    const double cs = cos(th), sn = sin(th);
    for (int i = 0; i < 5; ++i) {
        xy[2 * i] = d->local[i][0] * cs - d->local[i][1] * sn + x;
        xy[2 * i + 1] = d->local[i][0] * sn + d->local[i][1] * cs + y;
    }
    poly_from_world(p, 5, xy);
}

/* ShipGame.closest_goal (game.py:333-349) + the [-1,-1] default (ship_env.py:103-107) */
static void closest_goal(const double *goals, int alive, double x, double y, double *gx, double *gy, double *tie)
{
    double best = INFINITY, second = INFINITY;
    *gx = -1.0; *gy = -1.0;
    for (int k = 0; k < ORC_NGOALS; ++k) {
        if (!(alive >> k & 1)) continue;
        const double dx = goals[2 * k] - x, dy = goals[2 * k + 1] - y;
        const double dist = sqrt(dx * dx + dy * dy);
        if (dist < best) { second = best; best = dist; *gx = goals[2 * k]; *gy = goals[2 * k + 1]; }
        else if (dist < second) second = dist;
    }
    *tie = second - best;
}

/* ShipEnv.__add_states (ship_env.py:79-113): one 6+n_beams frame */
static void make_frame(const orc_config *c, const double *pose, int rudder, int alive, const double *goals,
                       const double *lidar, double *frame, double *tie)
{
    double gx, gy;
    closest_goal(goals, alive, pose[0], pose[1], &gx, &gy, tie);
    frame[0] = pose[0]; frame[1] = pose[1]; frame[2] = (double)rudder; frame[3] = pose[2];
    frame[4] = gx; frame[5] = gy;
    for (int i = 0; i < c->n_beams; ++i) frame[6 + i] = lidar[i];
}

static void push_frame(const orc_config *c, double *hist, const double *frame)
{
    const int F = ORC_FRAME(c->n_beams), H = c->history;
    memmove(hist, hist + F, sizeof(double) * (size_t)F * (H - 1));      /* deque(maxlen) .extend */
    memcpy(hist + (size_t)F * (H - 1), frame, sizeof(double) * F);
}

/* ShipEnv.reset (ship_env.py:171-184) + ShipGame.reset (game.py:260-277) for env e, scenario `scen` */
static void reset_env(const orc_config *c, const orc_bank *bank, const orc_state *s, int e, int scen, int episode)
{
    const int F = ORC_FRAME(c->n_beams);
    double *pose = s->pose + (size_t)e * 6;
    pose[0] = c->W / 2.0; pose[1] = c->spawn_y;                           /* game.py:274 */
    pose[2] = pose[3] = pose[4] = pose[5] = 0.0;
    memcpy(s->goals + (size_t)e * 10, bank->goals + (size_t)scen * 10, sizeof(double) * 10);
    for (int i = 0; i < ORC_MAXBEAMS; ++i) s->lidar[(size_t)e * ORC_MAXBEAMS + i] = -1.0;   /* models.py:36 */
    int32_t *in = s->ints + (size_t)e * 5;
    in[0] = 0; in[1] = (1 << ORC_NGOALS) - 1; in[2] = 0; in[3] = scen; in[4] = episode;
    s->ep_return[e] = 0.0;
    double *hist = s->hist + (size_t)e * F * c->history;
    for (int i = 0; i < F * c->history; ++i) hist[i] = -1.0;            /* ship_env.py:180-181 */
    double frame[ORC_FRAME(ORC_MAXBEAMS)], tie;
    make_frame(c, pose, 0, in[1], s->goals + (size_t)e * 10, s->lidar + (size_t)e * ORC_MAXBEAMS, frame, &tie);
    push_frame(c, hist, frame);
}

/* margins row layout: [0] |ship-bank SAT separation| (min over banks), [1] min_k |dist(goal_k)-r| over alive
 * goals, [2] out-of-bounds slack, [3] nearest-goal tie slack (new frame), [4+i] lidar ray i slack */
#define ORC_MARGIN_STRIDE (4 + 2 * ORC_MAXBEAMS)   /* [4+MAXBEAMS+i] = conditioning 1/|n.dir| of ray i's accepted hit (1 if none) */
/* flags bits */
#define ORC_F_COLLIDING 1
#define ORC_F_GOAL 2
#define ORC_F_OOB 4
#define ORC_F_TIMEOUT 8
#define ORC_F_ALLGOALS 16
/* stats layout (int64 / double pairs kept as double[16] for simplicity) */
enum { ST_EPISODES, ST_RETURN, ST_LENGTH, ST_GOALS, ST_COLLISION, ST_OOB, ST_TIMEOUT, ST_ALLGOALS, ST_STEPS, ST_N };

static void step_env(const orc_config *c, const orc_derived *d, const orc_bank *bank, const orc_state *s, int e,
                     int action, double *obs, double *reward_out, uint8_t *done_out, uint8_t *flags_out,
                     double *margins, double *stats)
{
    const int F = ORC_FRAME(c->n_beams);
    double *pose = s->pose + (size_t)e * 6;
    double *lidar = s->lidar + (size_t)e * ORC_MAXBEAMS;
    double *goals = s->goals + (size_t)e * 10;
    int32_t *in = s->ints + (size_t)e * 5;
    double *hist = s->hist + (size_t)e * F * c->history;
    const int scen = in[3];
    orc_poly bankp[2];
    load_bank_poly(c, bank, scen, 0, &bankp[0]);
    load_bank_poly(c, bank, scen, 1, &bankp[1]);
    double mg[ORC_MARGIN_STRIDE];
    for (int i = 0; i < ORC_MARGIN_STRIDE; ++i) mg[i] = INFINITY;
    for (int i = 0; i < ORC_MAXBEAMS; ++i) mg[4 + ORC_MAXBEAMS + i] = 1.0;

    /* -- ShipGame.handle_discrete_action (game.py:140-153), Ship.move_forward / rotate (models.py:129-146) */
    double fx = 0.0, fy = 0.0, torque = 0.0;
    if (action == 0) {
        /* cpBodyApplyForceAtLocalPoint((0,F), (-rudder, 0)); cog = (0,0) */
        const double cs = cos(pose[2]), sn = sin(pose[2]);
        fx = -c->thrust * sn;
        fy = c->thrust * cs;
        torque = (-(double)in[0]) * c->thrust;     /* cross((-rudder,0),(0,F)) */
    } else if (action == 1) {
        in[0] -= 5; if (in[0] < -10) in[0] = -10;
    } else if (action == 2) {
        in[0] += 5; if (in[0] > 10) in[0] = 10;
    }

    /* -- ShipGame.update: lidar BEFORE the physics step (game.py:193-194); LiDAR.query (models.py:39-76) */
    {
        orc_poly ship;
        ship_world(d, pose[0], pose[1], pose[2], &ship);
        double l = INFINITY, r = -INFINITY, b = INFINITY, t = -INFINITY;
        for (int i = 0; i < 5; ++i) { l = dmin(l, ship.vx[i]); r = dmax(r, ship.vx[i]); b = dmin(b, ship.vy[i]); t = dmax(t, ship.vy[i]); }
        const double ox = pose[0] + (r - l) / 2.0, oy = pose[1] + (t - b) / 2.0;     /* models.py:51-53 */
        const double cs = cos(pose[2]), sn = sin(pose[2]);
        for (int i = 0; i < c->n_beams; ++i) {
            /* cos(angle + a_i), sin(angle + a_i) */
            const double dc = cs * d->ray_cos[i] - sn * d->ray_sin[i];
            const double ds = sn * d->ray_cos[i] + cs * d->ray_sin[i];
            const double ex = ox + c->lidar_distance * dc, ey = oy + c->lidar_distance * ds;
            for (int k = 0; k < 2; ++k) {                 /* first shape in list order that reports a hit wins */
                orc_seg_info info;
                poly_segment_query(&bankp[k], ox, oy, ex, ey, 0.0, &info);
                mg[4 + i] = dmin(mg[4 + i], info.margin);
                if (info.hit) {
                    lidar[i] = sqrt((info.px - ox) * (info.px - ox) + (info.py - oy) * (info.py - oy));
                    mg[4 + ORC_MAXBEAMS + i] = info.cond;
                    break;
                }
            }
        }
    }

    /* -- cpSpaceStep (1): cpBodyUpdatePosition */
    const double dt = c->dt;
    pose[0] += pose[3] * dt;
    pose[1] += pose[4] * dt;
    pose[2] += pose[5] * dt;

    /* -- cpSpaceStep (2)+(3): overlap tests at the new pose; begin callbacks game.py:232-257 */
    int colliding = 0, goal_reached = 0;
    {
        orc_poly ship;
        ship_world(d, pose[0], pose[1], pose[2], &ship);
        for (int k = 0; k < 2; ++k) {
            double sep;
            if (polys_touch(&ship, &bankp[k], &sep)) colliding = 1;
            mg[0] = dmin(mg[0], fabs(sep));
        }
        for (int k = 0; k < ORC_NGOALS; ++k) {
            if (!(in[1] >> k & 1)) continue;
            const double dist = poly_point_distance(&ship, goals[2 * k], goals[2 * k + 1]);
            mg[1] = dmin(mg[1], fabs(dist - c->goal_radius));
            if (dist <= c->goal_radius) { goal_reached = 1; in[1] &= ~(1 << k); }
        }
    }

    /* -- cpSpaceStep (4): cpBodyUpdateVelocity; forces cleared */
    pose[3] = pose[3] * d->damping + (fx / c->mass) * dt;
    pose[4] = pose[4] * d->damping + (fy / c->mass) * dt;
    pose[5] = pose[5] * d->damping + (torque / d->moment) * dt;

    /* -- ShipEnv.determine_reward (ship_env.py:62-77) */
    const double x = pose[0], y = pose[1];
    const int oob = (x < 0.0 || x > c->W || y < 0.0 || y > c->H);
    mg[2] = dmin(dmin(fabs(x), fabs(c->W - x)), dmin(fabs(y), fabs(c->H - y)));
    double reward;
    if (goal_reached) reward = 1.0;
    else if (x < 0.0 || x > c->W) reward = -1.0;
    else if (y < 0.0 || y > c->H) reward = -1.0;
    else reward = c->step_penalty;
    s->ep_return[e] += reward;                                       /* ship_env.py:150 */

    /* -- __add_states (ship_env.py:151) */
    double frame[ORC_FRAME(ORC_MAXBEAMS)];
    make_frame(c, pose, in[0], in[1], goals, lidar, frame, &mg[3]);
    push_frame(c, hist, frame);
    in[2] += 1;                                                      /* ship_env.py:152 */

    /* -- is_done (ship_env.py:115-134) */
    const int all_goals = (in[1] == 0);
    const int timeout = (in[2] >= c->max_steps);
    const int done = colliding || all_goals || oob || timeout;
    uint8_t flags = (uint8_t)((colliding ? ORC_F_COLLIDING : 0) | (goal_reached ? ORC_F_GOAL : 0) | (oob ? ORC_F_OOB : 0)
                              | (timeout ? ORC_F_TIMEOUT : 0) | (all_goals ? ORC_F_ALLGOALS : 0));

    if (stats) {
        stats[ST_STEPS] += 1.0;
        if (goal_reached) stats[ST_GOALS] += 1.0;   /* steps on which a goal was reached */
    }
    if (done && stats) {
        stats[ST_EPISODES] += 1.0;
        stats[ST_RETURN] += s->ep_return[e];
        stats[ST_LENGTH] += (double)in[2];
        if (colliding) stats[ST_COLLISION] += 1.0;
        if (oob) stats[ST_OOB] += 1.0;
        if (timeout) stats[ST_TIMEOUT] += 1.0;
        if (all_goals) stats[ST_ALLGOALS] += 1.0;
    }
    if (done && c->auto_reset) {
        const int episode = in[4] + 1;
        const int ns = pick_for(c, c->env_id_offset + e, episode);
        reset_env(c, bank, s, e, ns, episode);
    }
    if (obs) memcpy(obs, hist, sizeof(double) * (size_t)F * c->history);
    if (reward_out) *reward_out = reward;
    if (done_out) *done_out = (uint8_t)done;
    if (flags_out) *flags_out = flags;
    if (margins) memcpy(margins, mg, sizeof(mg));
}

/* ------------------------------------------------------------------------------------------------
 * batched entry points
 * ---------------------------------------------------------------------------------------------- */
int orc_margin_stride(void) { return ORC_MARGIN_STRIDE; }
int orc_max_beams(void) { return ORC_MAXBEAMS; }
int orc_stats_len(void) { return ST_N; }
double orc_ship_moment(const orc_config *c) { orc_derived d; derive(c, &d); return d.moment; }
double orc_damping(const orc_config *c) { orc_derived d; derive(c, &d); return d.damping; }
void orc_ship_hull(const orc_config *c, double *xy10) { orc_derived d; derive(c, &d); memcpy(xy10, d.local, sizeof(double) * 10); }

/* reset envs whose mask byte is non-zero (mask NULL = all) to the given scenarios (scen NULL = philox pick
 * for episode `ints[e][4]+1`, or episode 0 when first != 0).  obs (optional) receives the reset obs. */
void orc_reset(const orc_config *c, const orc_bank *bank, const orc_state *s, int n, const uint8_t *mask,
               const int32_t *scen, int first, double *obs)
{
    const int F = ORC_FRAME(c->n_beams);
    for (int e = 0; e < n; ++e) {
        if (mask && !mask[e]) continue;
        const int episode = first ? 0 : s->ints[(size_t)e * 5 + 4] + 1;
        const int sc = scen ? scen[e] : pick_for(c, c->env_id_offset + e, episode);
        reset_env(c, bank, s, e, sc, episode);
        if (obs) memcpy(obs + (size_t)e * F * c->history, s->hist + (size_t)e * F * c->history, sizeof(double) * F * c->history);
    }
}

typedef struct {
    const orc_config *c; const orc_bank *bank; const orc_state *s;
    int e0, e1, n, K; const int32_t *actions;
    double *obs, *reward; uint8_t *done, *flags; double *margins; double stats[ST_N];
    uint32_t step0;
} orc_job;

static void *job_main(void *arg)
{
    orc_job *j = (orc_job *)arg;
    orc_derived d; derive(j->c, &d);
    const int F = ORC_FRAME(j->c->n_beams) * j->c->history;
    for (int k = 0; k < j->K; ++k)
        for (int e = j->e0; e < j->e1; ++e) {
            const size_t idx = (size_t)k * j->n + e;
            const int a = j->actions ? j->actions[idx] : orc_random_action(j->c->seed, j->c->env_id_offset + e, j->step0 + k);
            step_env(j->c, &d, j->bank, j->s, e, a,
                     j->obs ? j->obs + idx * F : NULL, j->reward ? j->reward + idx : NULL,
                     j->done ? j->done + idx : NULL, j->flags ? j->flags + idx : NULL,
                     j->margins ? j->margins + idx * ORC_MARGIN_STRIDE : NULL, j->stats);
        }
    return NULL;
}

/* K steps of all n envs.  actions[K][n] (NULL = philox random actions, step counter starting at step0);
 * outputs (each optional) obs[K][n][F*H], reward[K][n], done[K][n], flags[K][n], margins[K][n][stride];
 * stats[ST_N] is ACCUMULATED into.  n_threads > 1 splits the env range over pthreads (the
 * one-env-per-process pattern of train/stable_baselines/ppo.py:122-123, minus the pipes). */
void orc_step(const orc_config *c, const orc_bank *bank, const orc_state *s, int n, int K, const int32_t *actions,
              uint32_t step0, double *obs, double *reward, uint8_t *done, uint8_t *flags, double *margins,
              double *stats, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n;
    orc_job *jobs = (orc_job *)calloc((size_t)n_threads, sizeof(orc_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        orc_job *j = &jobs[t];
        j->c = c; j->bank = bank; j->s = s; j->n = n; j->K = K; j->actions = actions; j->step0 = step0;
        j->e0 = (int)((int64_t)n * t / n_threads); j->e1 = (int)((int64_t)n * (t + 1) / n_threads);
        j->obs = obs; j->reward = reward; j->done = done; j->flags = flags; j->margins = margins;
        if (n_threads == 1) job_main(j); else pthread_create(&th[t], NULL, job_main, j);
    }
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        if (stats) for (int i = 0; i < ST_N; ++i) stats[i] += jobs[t].stats[i];
    }
    free(jobs); free(th);
}
