// shipsim_device.cuh -- device-side data layout and per-env transition pieces for the fused step kernel.
//
// New code, written for sm_100a.  It implements the transition spec of SURVEY.md Appendix A, which restates
// ShipEnv.step (ship_gym/ship_env.py:136-156) and everything it reaches (game.py:140-153,185-195,232-257,
// 333-349; models.py:39-76,129-146) plus the Chipmunk2D routines the reference delegates to.  fp32 state,
// hull planes precomputed in double on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shipsim {

constexpr int kGoals = 5;
constexpr int kBeams = 10;
constexpr int kFrame = 16;
constexpr int kPlanes = 8;
constexpr int kShipVerts = 5;
constexpr int kStatSlots = 128;     // replicated accumulator rows (one 128-byte line each) to spread atomics
constexpr int kStatLen = 16;
constexpr int kBankHeader4 = 6;     // float4s before the edge records of a scenario

// Scenario record (float4 units): [0] aabb bank0 (l,b,r,t)  [1] aabb bank1  [2] goal0.xy goal1.xy
// [3] goal2.xy goal3.xy  [4] goal4.xy, bits(n0), bits(n1)  [5] reserved; then per bank b, edge i < n_b:
// [6 + b*maxv + i] = nx, ny, v_i.x, v_i.y
// Plane i is the edge v_{i-1} -> v_i with outward unit normal n (Chipmunk's cpSplittingPlane {v0 = v_i, n}).
// All plane tests are evaluated RELATIVE to the vertex (n.(p - v_i)), never as n.p - n.v_i: world coordinates
// are O(1000) and fp32 would lose ~1e-4 to cancellation; differences of nearby points are (nearly) exact.

struct StepParams {
    float4 *state;               // [kPlanes][N]
    const float4 *bank;          // packed scenario records
    const void *actions;         // [K][N]
    float4 *obs;                 // [K][N][4*history]
    float *reward;               // [K][N]
    uint8_t *done;               // [K][N]
    double *stats;               // [kStatSlots][kStatLen]
    unsigned long long seed;
    long long env_id_offset;
    int N, K;
    int action_dtype, history, auto_reset, max_steps;
    int n_scen, maxv, scen_stride4;
    unsigned step0;              // global step counter at launch (random-action stream)
    float W, H, dt, damping;
    float lidar_len;
    float acc_dt;                // thrust/mass*dt        (cpBodyUpdateVelocity: v += f*m_inv*dt)
    float ang_dt;                // thrust/moment*dt      (w += t*i_inv*dt, t = -rudder*thrust)
    float goal_r, step_penalty, spawn_x, spawn_y;
    float ray_c[kBeams], ray_s[kBeams];              // cos/sin of radians(90 - spread/2 + i*spread/n) (models.py:48-49,62)
    float ship_lx[kShipVerts], ship_ly[kShipVerts];  // body-frame hull, CCW (models.py:6,88 through cpConvexHull)
    float ship_nx[kShipVerts], ship_ny[kShipVerts];  // body-frame outward normal of edge j-1 -> j
    float ship_aabb[4];                              // body-frame l,b,r,t of the hull
};

struct EnvRegs {
    float x, y, th, vx, vy, w, ret;
    int rudder, alive, steps, scen, episode;
    float lid[kBeams];
    float g[2 * kGoals];
};

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long seed, unsigned long long ctr_lo, unsigned c2, unsigned c3)
{
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    unsigned c0 = (unsigned)ctr_lo, c1 = (unsigned)(ctr_lo >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ int pick_scenario(const StepParams &p, long long gid, int episode)
{
    const uint4 r = philox4x32_10(p.seed, (unsigned long long)gid, (unsigned)episode, 0u);
    return (int)__umulhi(r.x, (unsigned)p.n_scen);
}

__device__ __forceinline__ int random_action(const StepParams &p, long long gid, unsigned step)
{
    const uint4 r = philox4x32_10(p.seed, (unsigned long long)gid, step, 1u);
    return (int)__umulhi(r.x, 3u);
}

// ---------------------------------------------------------------------------------------------- state I/O
__device__ __forceinline__ int pack_bits(int rudder, int alive, int steps)
{
    return ((rudder / 5 + 2) & 7) | ((alive & 31) << 3) | (steps << 8);
}

__device__ __forceinline__ void load_env(const StepParams &p, int e, EnvRegs &r)
{
    const float4 *s = p.state;
    const size_t N = (size_t)p.N;
    const float4 a = s[0 * N + e], b = s[1 * N + e], l0 = s[2 * N + e], l1 = s[3 * N + e], l2 = s[4 * N + e];
    const float4 g0 = s[5 * N + e], g1 = s[6 * N + e], g2 = s[7 * N + e];
    r.x = a.x; r.y = a.y; r.th = a.z; r.vx = a.w;
    r.vy = b.x; r.w = b.y; r.ret = b.z;
    const int bits = __float_as_int(b.w);
    r.rudder = ((bits & 7) - 2) * 5; r.alive = (bits >> 3) & 31; r.steps = bits >> 8;
    r.lid[0] = l0.x; r.lid[1] = l0.y; r.lid[2] = l0.z; r.lid[3] = l0.w;
    r.lid[4] = l1.x; r.lid[5] = l1.y; r.lid[6] = l1.z; r.lid[7] = l1.w;
    r.lid[8] = l2.x; r.lid[9] = l2.y; r.scen = __float_as_int(l2.z); r.episode = __float_as_int(l2.w);
    r.g[0] = g0.x; r.g[1] = g0.y; r.g[2] = g0.z; r.g[3] = g0.w;
    r.g[4] = g1.x; r.g[5] = g1.y; r.g[6] = g1.z; r.g[7] = g1.w;
    r.g[8] = g2.x; r.g[9] = g2.y;
}

__device__ __forceinline__ void store_env(const StepParams &p, int e, const EnvRegs &r, bool goals_dirty)
{
    float4 *s = p.state;
    const size_t N = (size_t)p.N;
    s[0 * N + e] = make_float4(r.x, r.y, r.th, r.vx);
    s[1 * N + e] = make_float4(r.vy, r.w, r.ret, __int_as_float(pack_bits(r.rudder, r.alive, r.steps)));
    s[2 * N + e] = make_float4(r.lid[0], r.lid[1], r.lid[2], r.lid[3]);
    s[3 * N + e] = make_float4(r.lid[4], r.lid[5], r.lid[6], r.lid[7]);
    s[4 * N + e] = make_float4(r.lid[8], r.lid[9], __int_as_float(r.scen), __int_as_float(r.episode));
    if (goals_dirty) {           // goals only change on reset
        s[5 * N + e] = make_float4(r.g[0], r.g[1], r.g[2], r.g[3]);
        s[6 * N + e] = make_float4(r.g[4], r.g[5], r.g[6], r.g[7]);
        s[7 * N + e] = make_float4(r.g[8], r.g[9], 0.f, 0.f);
    }
}

// ShipGame.reset + ShipEnv.reset for one env (game.py:260-277, ship_env.py:171-184); goals come from the bank.
__device__ __forceinline__ void reset_env(const StepParams &p, EnvRegs &r, int scen, int episode)
{
    const float4 *sc = p.bank + (size_t)scen * p.scen_stride4;
    const float4 g0 = __ldg(sc + 2), g1 = __ldg(sc + 3), g2 = __ldg(sc + 4);
    r.g[0] = g0.x; r.g[1] = g0.y; r.g[2] = g0.z; r.g[3] = g0.w;
    r.g[4] = g1.x; r.g[5] = g1.y; r.g[6] = g1.z; r.g[7] = g1.w;
    r.g[8] = g2.x; r.g[9] = g2.y;
    r.x = p.spawn_x; r.y = p.spawn_y; r.th = 0.f; r.vx = 0.f; r.vy = 0.f; r.w = 0.f; r.ret = 0.f;
    r.rudder = 0; r.alive = (1 << kGoals) - 1; r.steps = 0; r.scen = scen; r.episode = episode;
#pragma unroll
    for (int i = 0; i < kBeams; ++i) r.lid[i] = -1.f;         // models.py:36
}

// ShipGame.closest_goal (game.py:333-349): first strict minimum wins; (-1,-1) when none (ship_env.py:103-107)
__device__ __forceinline__ void closest_goal(const EnvRegs &r, float &gx, float &gy)
{
    float best = 3.0e38f;
    gx = -1.f; gy = -1.f;
#pragma unroll
    for (int k = 0; k < kGoals; ++k) {
        const float dx = r.g[2 * k] - r.x, dy = r.g[2 * k + 1] - r.y;
        const float d2 = dx * dx + dy * dy;
        const bool take = ((r.alive >> k) & 1) && d2 < best;
        if (take) { best = d2; gx = r.g[2 * k]; gy = r.g[2 * k + 1]; }
    }
}

// ---------------------------------------------------------------------------------------------- lidar
// LiDAR.query (models.py:39-76) over cpShapeSegmentQuery / cpPolyShapeSegmentQuery (Chipmunk, r = 0).
// Sampled at the PRE-integration pose (game.py:193 before :194).  Hits overwrite r.lid[i]; misses keep the
// previous value (sticky vals, models.py:71).  First bank in list order that reports a hit wins (models.py:61-72).
__device__ __forceinline__ void lidar_query(const StepParams &p, const float4 *__restrict__ sc, EnvRegs &r, float c, float s)
{
    // cached AABB of the rotated hull -> ray origin = body origin + half extents (models.py:51-53)
    float minx = 0.f, maxx = 0.f, miny = 0.f, maxy = 0.f;      // hull vertex 0 is the body origin
#pragma unroll
    for (int j = 1; j < kShipVerts; ++j) {
        const float wx = p.ship_lx[j] * c - p.ship_ly[j] * s;
        const float wy = p.ship_lx[j] * s + p.ship_ly[j] * c;
        minx = fminf(minx, wx); maxx = fmaxf(maxx, wx); miny = fminf(miny, wy); maxy = fmaxf(maxy, wy);
    }
    const float hx = 0.5f * (maxx - minx), hy = 0.5f * (maxy - miny);
    const float ox = r.x + hx, oy = r.y + hy;
    const float L = p.lidar_len;
    const float4 hdr = __ldg(sc + 4);
    unsigned pending = (1u << kBeams) - 1u;
#pragma unroll 1
    for (int b = 0; b < 2; ++b) {
        const float4 bb = __ldg(sc + b);
        if (ox + L < bb.x || ox - L > bb.z || oy + L < bb.y || oy - L > bb.w) continue;   // fan cannot reach the bank
        const float4 *E = sc + kBankHeader4 + b * p.maxv;
        const int n = __float_as_int(b == 0 ? hdr.z : hdr.w);
        // pass 1: cpShapePointQuery "inside" test + edges whose plane is within reach in front of the origin
        unsigned live = 0u;
        bool inside = true;
        for (int i = 0; i < n; ++i) {
            const float4 e = __ldg(E + i);
            const float d = e.x * ((r.x - e.z) + hx) + e.y * ((r.y - e.w) + hy);      // n.(origin - v_i)
            inside = inside && (d <= 0.f);
            if (d >= 0.f && d <= L) live |= 1u << i;
        }
        if (inside) {                    // start point inside the shape: alpha = 0, point stays at the ray end
#pragma unroll
            for (int i = 0; i < kBeams; ++i) if (pending >> i & 1u) r.lid[i] = L;
            pending = 0u;
            break;
        }
        unsigned hit = 0u;
        while (live) {
            const int i = __ffs(live) - 1;
            live &= live - 1u;
            const float4 e = __ldg(E + i);
            const float4 ep = __ldg(E + (i == 0 ? n - 1 : i - 1));
            // The hit distance is d / (-n.dir): any error of d is amplified by 1/cos(incidence).  The stored fp32
            // normal is only good to ~6e-8 rad, which over a 1000-unit edge is 6e-5 of d -- so for the (few) live
            // edges the normal is rebuilt in double from the two fp32 vertices (exactly what the reference's
            // double-precision planes are made of) and d is formed in double.  B200 runs FP64 at half FP32 rate.
            const double exd = (double)e.z - (double)ep.z, eyd = (double)e.w - (double)ep.w;
            const double inv = rsqrt(exd * exd + eyd * eyd);
            const double nxd = eyd * inv, nyd = -exd * inv;
            const double qxd = ((double)r.x - (double)e.z) + (double)hx, qyd = ((double)r.y - (double)e.w) + (double)hy;
            const float d = (float)(nxd * qxd + nyd * qyd);
            const float ta = (float)(nxd * qyd - nyd * qxd);                          // cross(n, origin - v_i)
            const float tmin = -(float)((exd * exd + eyd * eyd) * inv);               // cross(n, v_{i-1} - v_i) = -|edge|
            const float enx = (float)nxd, eny = (float)nyd;
#pragma unroll
            for (int k = 0; k < kBeams; ++k) {
                const float dx = c * p.ray_c[k] - s * p.ray_s[k];     // cos(angle + a_k)
                const float dy = s * p.ray_c[k] + c * p.ray_s[k];
                const float denom = -L * (enx * dx + eny * dy);       // an - bn
                float t;
                if (denom > 0.f) t = d / denom; else t = (d == 0.f) ? 0.f : 2.f;   // d / max(an-bn, DBL_MIN)
                const float tang = ta + t * L * (enx * dy - eny * dx);             // cross(n, lerp(a,b,t) - v_i)
                const bool ok = (t <= 1.f) && (tang >= tmin) && (tang <= 0.f) && (pending >> k & 1u);
                if (ok) { r.lid[k] = t * L; hit |= 1u << k; }
            }
        }
        pending &= ~hit;
    }
}

// ---------------------------------------------------------------------------------------------- overlap tests
// cpSpaceStep narrow phase at the post-integration pose.  Poly-vs-poly contact <=> no separating axis among
// the edge normals of both convex polygons (touching counts: GJK distance <= 0).
__device__ __forceinline__ bool ship_touches_bank(const StepParams &p, const float4 *__restrict__ sc, int b, int n,
                                                  float x, float y, const float (&rx)[kShipVerts], const float (&ry)[kShipVerts],
                                                  float c, float s, float sminx, float sminy, float smaxx, float smaxy)
{
    // rx, ry: hull vertices relative to the body origin (x, y); s* : world AABB of the hull
    const float4 bb = __ldg(sc + b);
    if (sminx > bb.z || smaxx < bb.x || sminy > bb.w || smaxy < bb.y) return false;     // cpBBIntersects (inclusive)
    const float4 *E = sc + kBankHeader4 + b * p.maxv;
    for (int i = 0; i < n; ++i) {                    // bank edge normals
        const float4 e = __ldg(E + i);
        const float base = e.x * (x - e.z) + e.y * (y - e.w);
        float m = e.x * rx[0] + e.y * ry[0];
#pragma unroll
        for (int k = 1; k < kShipVerts; ++k) m = fminf(m, e.x * rx[k] + e.y * ry[k]);
        if (base + m > 0.f) return false;
    }
#pragma unroll 1
    for (int j = 0; j < kShipVerts; ++j) {           // ship edge normals
        const float nx = p.ship_nx[j] * c - p.ship_ny[j] * s;
        const float ny = p.ship_nx[j] * s + p.ship_ny[j] * c;
        const float off = nx * rx[j] + ny * ry[j];
        float m = 3.0e38f;
        for (int i = 0; i < n; ++i) {
            const float4 e = __ldg(E + i);
            m = fminf(m, nx * (e.z - x) + ny * (e.w - y));
        }
        if (m - off > 0.f) return false;
    }
    return true;
}

// Circle-vs-poly contact (CircleToPoly): distance(goal centre, ship polygon) <= goal radius, evaluated in the
// body frame where the hull is constant.  cpPolyShapePointQuery semantics: inside => negative distance.
__device__ __forceinline__ bool goal_touches_ship(const StepParams &p, float qx, float qy)
{
    const float rr = p.goal_r;
    if (qx < p.ship_aabb[0] - rr || qx > p.ship_aabb[2] + rr || qy < p.ship_aabb[1] - rr || qy > p.ship_aabb[3] + rr)
        return false;
    bool outside = false;
    float best = 3.0e38f;
#pragma unroll
    for (int j = 0; j < kShipVerts; ++j) {
        const int j0 = (j + kShipVerts - 1) % kShipVerts;
        const float ax = p.ship_lx[j0], ay = p.ship_ly[j0], bx = p.ship_lx[j], by = p.ship_ly[j];
        outside = outside || (p.ship_nx[j] * (qx - bx) + p.ship_ny[j] * (qy - by) > 0.f);
        const float ex = ax - bx, ey = ay - by;                       // cpClosetPointOnSegment
        float t = (ex * (qx - bx) + ey * (qy - by)) / (ex * ex + ey * ey);
        t = fminf(fmaxf(t, 0.f), 1.f);
        const float cx = bx + ex * t - qx, cy = by + ey * t - qy;
        best = fminf(best, cx * cx + cy * cy);
    }
    return !outside || best <= rr * rr;
}

}  // namespace shipsim
