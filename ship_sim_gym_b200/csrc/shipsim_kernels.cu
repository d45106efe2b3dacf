// shipsim_kernels.cu -- the fused ShipEnv step kernel (K env-steps per launch), reset and stats kernels.
// sm_100a only.  See shipsim_device.cuh for the data layout, the execution model (G lanes per env, warp-wide
// cooperative geometry) and the reference lines each piece restates.
#include "shipsim_device.cuh"
#include "shipsim_launch.h"

namespace shipsim {

__device__ __forceinline__ int load_action(const StepParams &p, int k, int e, long long gid)
{
    const size_t idx = (size_t)k * p.N + e;
    switch (p.action_dtype) {
        case 0: return __ldg(reinterpret_cast<const int *>(p.actions) + idx);
        case 1: return (int)__ldg(reinterpret_cast<const long long *>(p.actions) + idx);
        case 2: return (int)__ldg(reinterpret_cast<const unsigned char *>(p.actions) + idx);
        default: return random_action(p, gid, p.step0 + (unsigned)k);
    }
}

struct ScenConsts { float4 bb0, bb1; int n0, n1; };

__device__ __forceinline__ void load_scen_consts(const StepParams &p, int scen, ScenConsts &sc)
{
    const float4 *rec = p.bank + (size_t)scen * p.scen_stride4;
    sc.bb0 = __ldg(rec + 0);
    sc.bb1 = __ldg(rec + 1);
    const float4 h = __ldg(rec + 4);
    sc.n0 = __float_as_int(h.z);
    sc.n1 = __float_as_int(h.w);
}

// extents of the rotated hull relative to the body origin: the shape's cached AABB (cpPolyShapeCacheData)
__device__ __forceinline__ void hull_extents(const StepParams &p, float c, float s, float &minx, float &maxx, float &miny, float &maxy)
{
    minx = 0.f; maxx = 0.f; miny = 0.f; maxy = 0.f;             // hull vertex 0 is the body origin
#pragma unroll
    for (int j = 1; j < kShipVerts; ++j) {
        const float wx = p.ship_lx[j] * c - p.ship_ly[j] * s;
        const float wy = p.ship_lx[j] * s + p.ship_ly[j] * c;
        minx = fminf(minx, wx); maxx = fmaxf(maxx, wx); miny = fminf(miny, wy); maxy = fmaxf(maxy, wy);
    }
}

// ------------------------------------------------------------------------------------------------------------
// G lanes per env, 32/G envs per warp.  Scalar state is loaded once, lives in registers for K steps and is stored
// once; the two most recent observation frames of every env -- including the sticky lidar readings, which are
// state -- live in a padded shared-memory tile from which each step's obs rows are copied out, fully coalesced.
// ------------------------------------------------------------------------------------------------------------
template <int G, int HIST>
__global__ void __launch_bounds__(kThreads) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int EPW = 32 / G;                 // envs per warp
    constexpr int OBS4 = 4 * HIST;              // float4 per obs row
    constexpr int ROW4 = OBS4 + 1;              // padded tile row (odd float4 stride: conflict-free 128-bit accesses)
    constexpr int CF = 16 * (HIST - 1);         // float offset of the newest frame inside a row
    constexpr unsigned GMASK = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    __shared__ float4 s_tile[(kThreads / 32) * EPW * ROW4];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int gshift = grp * G;
    const int warp_env0 = (blockIdx.x * (kThreads / 32) + warp) * EPW;
    const int e = warp_env0 + grp;
    const bool valid = e < p.N;
    const bool leader = valid && gl == 0;
    float4 *tile = s_tile + warp * EPW * ROW4;
    float4 *row4 = tile + grp * ROW4;                               // this env's [older frame | newest frame]
    float *lidf = reinterpret_cast<float *>(row4) + CF + 6;         // newest frame's lidar slots = the sticky vals
    const float L = p.lidar_len;

    float st_episodes = 0.f, st_return = 0.f, st_length = 0.f, st_goal = 0.f;
    float st_coll = 0.f, st_oob = 0.f, st_timeout = 0.f, st_allgoals = 0.f;

    EnvRegs r;
    load_env(p, valid ? e : p.N - 1, r);        // lanes of idle groups shadow the last env; they never store
    ScenConsts sc;
    load_scen_consts(p, r.scen, sc);
    const long long gid = p.env_id_offset + e;
    float c, s;
    sincos_fast(r.th, s, c);
    float gx, gy;
    closest_goal(r, gx, gy);
    bool goals_dirty = false;
    // extents of the rotated hull relative to the body origin (cached AABB of the shape)
    float hminx, hmaxx, hminy, hmaxy;
    hull_extents(p, c, s, hminx, hmaxx, hminy, hmaxy);
    if (gl == 0) {                               // newest frame of the resident tile = frame of the current state
        row4[OBS4 - 4] = make_float4(r.x, r.y, (float)r.rudder, r.th);
        row4[OBS4 - 3] = make_float4(gx, gy, r.lid[0], r.lid[1]);
        row4[OBS4 - 2] = make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]);
        row4[OBS4 - 1] = make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]);
    }
    int a_next = load_action(p, 0, valid ? e : p.N - 1, gid);
    __syncwarp();

#pragma unroll 1
    for (int k = 0; k < p.K; ++k) {
        const int a = a_next;
        if (k + 1 < p.K) a_next = load_action(p, k + 1, valid ? e : p.N - 1, gid);    // prefetch: off the critical path
        const float4 *rec = p.bank + (size_t)r.scen * p.scen_stride4;
        const float4 *E0 = rec + kBankHeader4, *E1 = E0 + p.maxv;
        // previous frame <- newest frame of the last step / reset (SURVEY.md App. A note N2); lidar stays in place
        if (HIST == 2 && gl == 0) { row4[0] = row4[4]; row4[1] = row4[5]; row4[2] = row4[6]; row4[3] = row4[7]; }
        if (HIST == 2 && G > 1) __syncwarp();      // the copy has read the old readings before any lane overwrites them

        // ---- ShipGame.handle_discrete_action (game.py:140-153); Ship.move_forward / rotate (models.py:129-146)
        float dvx = 0.f, dvy = 0.f, dw = 0.f;
        if (a == 0) { dvx = -p.acc_dt * s; dvy = p.acc_dt * c; dw = -p.ang_dt * (float)r.rudder; }
        else if (a == 1) r.rudder = max(r.rudder - 5, -10);
        else if (a == 2) r.rudder = min(r.rudder + 5, 10);

        // ---- LiDAR.query (models.py:39-76) at the PRE-integration pose (game.py:193 precedes :194)
        // ray origin = body origin + half the extents of the hull's cached AABB (models.py:51-53)
        const float hx = 0.5f * (hmaxx - hminx), hy = 0.5f * (hmaxy - hminy);
        // box of the fan relative to the origin (centre fcx,fcy; half sizes fhw,fhh): sector between the first and
        // the last ray, which contains every ray
        float fcx, fcy, fhw, fhh;
        {
            float x0 = -1.f, x1 = 1.f, y0 = -1.f, y1 = 1.f;
            if (p.fan_is_sector) {
                const float ax = c * p.ray_c[0] - s * p.ray_s[0], ay = s * p.ray_c[0] + c * p.ray_s[0];
                const float bx = c * p.ray_c[kBeams - 1] - s * p.ray_s[kBeams - 1], by = s * p.ray_c[kBeams - 1] + c * p.ray_s[kBeams - 1];
                x1 = (ay <= 0.f && by >= 0.f) ? 1.f : fmaxf(0.f, fmaxf(ax, bx));
                x0 = (ay >= 0.f && by <= 0.f) ? -1.f : fminf(0.f, fminf(ax, bx));
                y1 = (ax >= 0.f && bx <= 0.f) ? 1.f : fmaxf(0.f, fmaxf(ay, by));
                y0 = (ax <= 0.f && bx >= 0.f) ? -1.f : fminf(0.f, fminf(ay, by));
            }
            fcx = 0.5f * L * (x0 + x1); fhw = 0.5f * L * (x1 - x0) + 1e-3f;
            fcy = 0.5f * L * (y0 + y1); fhh = 0.5f * L * (y1 - y0) + 1e-3f;
        }
        bool reach0, reach1;
        {
            const float ox = r.x + hx + fcx, oy = r.y + hy + fcy;       // fan box centre, world
            reach0 = valid && !(ox + fhw < sc.bb0.x || ox - fhw > sc.bb0.z || oy + fhh < sc.bb0.y || oy - fhh > sc.bb0.w);
            reach1 = valid && !(ox + fhw < sc.bb1.x || ox - fhw > sc.bb1.z || oy + fhh < sc.bb1.y || oy - fhh > sc.bb1.w);
        }
        // stage A (cpShapePointQuery + plane culling), edges strided over the group: is the origin inside the bank,
        // and which planes face the origin and cut the fan box?
        unsigned cand0 = 0u, cand1 = 0u, out0 = 0u, out1 = 0u;
        if (G == 1) {
            if (reach0)
                for (int i = 0; i < sc.n0; ++i) {
                    const float4 ed = __ldg(E0 + i);
                    const float d = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);     // n.(origin - v_i)
                    const float dmin = d + (ed.x * fcx + ed.y * fcy) - (fabsf(ed.x) * fhw + fabsf(ed.y) * fhh);
                    out0 |= (d > 0.f) ? 1u : 0u;
                    cand0 |= (d >= 0.f && dmin <= 0.f) ? (1u << i) : 0u;
                }
            if (reach1)
                for (int i = 0; i < sc.n1; ++i) {
                    const float4 ed = __ldg(E1 + i);
                    const float d = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);
                    const float dmin = d + (ed.x * fcx + ed.y * fcy) - (fabsf(ed.x) * fhw + fabsf(ed.y) * fhh);
                    out1 |= (d > 0.f) ? 1u : 0u;
                    cand1 |= (d >= 0.f && dmin <= 0.f) ? (1u << i) : 0u;
                }
        } else {
            for (int it = 0; it * G < p.maxv; ++it) {                  // warp-uniform trip count
                const int i = it * G + gl;
                const bool a0 = reach0 && i < sc.n0, a1 = reach1 && i < sc.n1;
                float d0 = -1.f, d1 = -1.f, m0 = 1.f, m1 = 1.f;
                if (a0) {
                    const float4 ed = __ldg(E0 + i);
                    d0 = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);
                    m0 = d0 + (ed.x * fcx + ed.y * fcy) - (fabsf(ed.x) * fhw + fabsf(ed.y) * fhh);
                }
                if (a1) {
                    const float4 ed = __ldg(E1 + i);
                    d1 = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);
                    m1 = d1 + (ed.x * fcx + ed.y * fcy) - (fabsf(ed.x) * fhw + fabsf(ed.y) * fhh);
                }
                const unsigned bl0 = __ballot_sync(kFull, a0 && d0 >= 0.f && m0 <= 0.f), bo0 = __ballot_sync(kFull, a0 && d0 > 0.f);
                const unsigned bl1 = __ballot_sync(kFull, a1 && d1 >= 0.f && m1 <= 0.f), bo1 = __ballot_sync(kFull, a1 && d1 > 0.f);
                cand0 |= ((bl0 >> gshift) & GMASK) << (it * G); out0 |= (bo0 >> gshift) & GMASK;
                cand1 |= ((bl1 >> gshift) & GMASK) << (it * G); out1 |= (bo1 >> gshift) & GMASK;
            }
        }
        // stage B + rays, per group with the rays strided over its lanes (lane gl owns rays gl, gl+G, ...): no
        // cross-lane traffic, hits go straight into the resident frame.  cpPolyShapeSegmentQuery: later edges
        // overwrite earlier ones; LiDAR.query: the first bank (list order) that reports a hit wins, misses keep
        // the old reading (sticky vals, models.py:71).
        if ((reach0 && (out0 == 0u || cand0)) || (reach1 && (out1 == 0u || cand1))) {
            unsigned pend = 0u;
#pragma unroll
            for (int j = gl; j < kBeams; j += G) pend |= 1u << j;        // my rays
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const bool rb = b == 0 ? reach0 : reach1;
                if (!rb || pend == 0u) continue;
                const unsigned outb = b == 0 ? out0 : out1;
                unsigned cand = b == 0 ? cand0 : cand1;
                const int nb = b == 0 ? sc.n0 : sc.n1;
                const float4 *E = b == 0 ? E0 : E1;
                if (outb == 0u) {            // start point inside the shape: alpha = 0, point stays at the ray end
#pragma unroll
                    for (int j = gl; j < kBeams; j += G) if (pend >> j & 1u) lidf[j] = L;
                    pend = 0u;
                    continue;
                }
                unsigned hitm = 0u;
                while (cand) {
                    const int i = __ffs(cand) - 1;
                    cand &= cand - 1u;
                    const float4 ed = __ldg(E + i);
                    const float4 ep = __ldg(E + (i == 0 ? nb - 1 : i - 1));
                    {   // stage B (fp32): can the fan box reach the edge's extent along the plane?
                        const float qx = (r.x - ed.z) + hx, qy = (r.y - ed.w) + hy;
                        const float tc = (ed.x * qy - ed.y * qx) + (ed.x * fcy - ed.y * fcx);
                        const float te = fabsf(ed.y) * fhw + fabsf(ed.x) * fhh;
                        const float tmin = ed.x * (ep.w - ed.w) - ed.y * (ep.z - ed.z);
                        if (tc + te < tmin || tc - te > 0.f) continue;
                    }
                    // The hit distance is d / (-n.dir): an error of d is amplified by 1/cos(incidence).  The stored
                    // fp32 normal is good to ~6e-8 rad, i.e. 6e-5 of d over a 1000-unit edge, so for live edges
                    // the plane is rebuilt in double from the two fp32 vertices (what the reference's double
                    // planes are made of).  FP64 runs at half the FP32 rate on B200 and few lane-steps get here.
                    const double exd = (double)ed.z - (double)ep.z, eyd = (double)ed.w - (double)ep.w;
                    const double len2 = exd * exd + eyd * eyd;
                    const double inv = rsqrt(len2);
                    const double nxd = eyd * inv, nyd = -exd * inv;
                    const double qxd = ((double)r.x - (double)ed.z) + (double)hx;
                    const double qyd = ((double)r.y - (double)ed.w) + (double)hy;
                    const float d = (float)(nxd * qxd + nyd * qyd);
                    const float ta = (float)(nxd * qyd - nyd * qxd);       // cross(n, origin - v_i)
                    const float tmin = -(float)(len2 * inv);               // cross(n, v_{i-1} - v_i) = -|edge|
                    const float enx = (float)nxd, eny = (float)nyd;
                    if (d < 0.f) continue;
#pragma unroll
                    for (int j = gl; j < kBeams; j += G) {
                        const float dirx = c * p.ray_c[j] - s * p.ray_s[j], diry = s * p.ray_c[j] + c * p.ray_s[j];
                        const float denom = -L * (enx * dirx + eny * diry);    // an - bn
                        float t;
                        if (denom > 0.f) t = __fdividef(d, denom); else t = (d == 0.f) ? 0.f : 2.f;   // d / max(an-bn, DBL_MIN)
                        const float tang = ta + t * L * (enx * diry - eny * dirx);                  // cross(n, hit - v_i)
                        if ((pend >> j & 1u) && (t <= 1.f) && (tang >= tmin) && (tang <= 0.f)) { lidf[j] = t * L; hitm |= 1u << j; }
                    }
                }
                pend &= ~hitm;
            }
        }

        // ---- cpSpaceStep: positions first (cpBodyUpdatePosition)
        r.x += r.vx * p.dt;
        r.y += r.vy * p.dt;
        r.th += r.w * p.dt;
        sincos_fast(r.th, s, c);
        hull_extents(p, c, s, hminx, hmaxx, hminy, hmaxy);

        // ---- overlap tests at the new pose -> begin callbacks collide_ship / collide_goal (game.py:232-257)
        // cpBBIntersects (inclusive) pre-filter of the narrow phase
        const bool ov0 = valid && !(r.x + hminx > sc.bb0.z || r.x + hmaxx < sc.bb0.x || r.y + hminy > sc.bb0.w || r.y + hmaxy < sc.bb0.y);
        const bool ov1 = valid && !(r.x + hminx > sc.bb1.z || r.x + hmaxx < sc.bb1.x || r.y + hminy > sc.bb1.w || r.y + hmaxy < sc.bb1.y);
        bool colliding = false;
        {
            // cooperative separating-axis test: lanes <-> bank edges (n <= 32).  Contact <=> no separating axis
            // among the edge normals of both convex polygons (touching counts: GJK distance <= 0).
            unsigned needy = __ballot_sync(kFull, gl == 0 && (ov0 || ov1));
            while (needy) {
                const int src = __ffs(needy) - 1;
                needy &= needy - 1u;
                const float bx = __shfl_sync(kFull, r.x, src), by = __shfl_sync(kFull, r.y, src);
                const float bc = __shfl_sync(kFull, c, src), bs = __shfl_sync(kFull, s, src);
                const int bscen = __shfl_sync(kFull, r.scen, src);
                const int bflags = __shfl_sync(kFull, sc.n0 | (sc.n1 << 8) | ((int)ov0 << 16) | ((int)ov1 << 17), src);
                const float4 *bE = p.bank + (size_t)bscen * p.scen_stride4 + kBankHeader4;
                float rx[kShipVerts], ry[kShipVerts];
#pragma unroll
                for (int j = 0; j < kShipVerts; ++j) {
                    rx[j] = p.ship_lx[j] * bc - p.ship_ly[j] * bs;
                    ry[j] = p.ship_lx[j] * bs + p.ship_ly[j] * bc;
                }
                bool coll = false;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    if (((bflags >> (16 + b)) & 1) && !coll) {
                        const int nb = (bflags >> (8 * b)) & 0xff;
                        const bool act = lane < nb;
                        float4 ed = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (act) ed = __ldg(bE + b * p.maxv + lane);
                        float m = ed.x * rx[0] + ed.y * ry[0];
#pragma unroll
                        for (int j = 1; j < kShipVerts; ++j) m = fminf(m, ed.x * rx[j] + ed.y * ry[j]);
                        const float base = ed.x * (bx - ed.z) + ed.y * (by - ed.w);
                        bool sep = __ballot_sync(kFull, act && base + m > 0.f) != 0u;      // a bank edge normal separates
                        if (!sep) {
#pragma unroll
                            for (int j = 0; j < kShipVerts; ++j) {                          // ship edge normals
                                const float nx = p.ship_nx[j] * bc - p.ship_ny[j] * bs;
                                const float ny = p.ship_nx[j] * bs + p.ship_ny[j] * bc;
                                const float off = nx * rx[j] + ny * ry[j];
                                const float pr = act ? nx * (ed.z - bx) + ny * (ed.w - by) : 3.0e38f;
                                const int mn = __reduce_min_sync(kFull, f2ord(pr));
                                sep = sep || (mn > f2ord(off));
                            }
                        }
                        coll = !sep;
                    }
                }
                if (coll && (lane / G) == (src / G)) colliding = true;
            }
        }
        bool goal_reached = false;
        float gd2[kGoals];
        {
            // goals: squared distance to the body origin serves both the closest-goal search and a bounding-circle
            // cull; only goals inside the circle are rotated into the body frame for the exact distance test
            unsigned cand = 0u;
#pragma unroll
            for (int g = 0; g < kGoals; ++g) {
                const float ux = r.g[2 * g] - r.x, uy = r.g[2 * g + 1] - r.y;
                gd2[g] = ux * ux + uy * uy;
                if (valid && ((r.alive >> g) & 1) && gd2[g] <= p.goal_cull_r2) cand |= 1u << g;
            }
            while (cand) {
                const int g = __ffs(cand) - 1;
                cand &= cand - 1u;
                float ux = r.g[0], uy = r.g[1];
#pragma unroll
                for (int j = 1; j < kGoals; ++j) if (g == j) { ux = r.g[2 * j]; uy = r.g[2 * j + 1]; }
                ux -= r.x; uy -= r.y;
                const float qx = ux * c + uy * s, qy = -ux * s + uy * c;
                if (!goal_culled(p, qx, qy) && goal_touches_ship(p, qx, qy)) { goal_reached = true; r.alive &= ~(1 << g); }
            }
        }

        // ---- cpBodyUpdateVelocity: v = v*damping + f/m*dt, w = w*damping + t/I*dt
        r.vx = r.vx * p.damping + dvx;
        r.vy = r.vy * p.damping + dvy;
        r.w = r.w * p.damping + dw;

        // ---- ShipEnv.determine_reward (ship_env.py:62-77): collision alone does not change the value (Q12)
        const bool oob = (r.x < 0.f) || (r.x > p.W) || (r.y < 0.f) || (r.y > p.H);
        const float reward = goal_reached ? 1.f : (oob ? -1.f : p.step_penalty);
        r.ret += reward;
        r.steps += 1;
        {   // ShipGame.closest_goal (game.py:333-349) over the goals still alive
            float best = 3.0e38f;
            gx = -1.f; gy = -1.f;
#pragma unroll
            for (int g = 0; g < kGoals; ++g)
                if (((r.alive >> g) & 1) && gd2[g] < best) { best = gd2[g]; gx = r.g[2 * g]; gy = r.g[2 * g + 1]; }
        }
        const bool all_goals = (r.alive == 0);
        const bool timeout = (r.steps >= p.max_steps);
        const bool done = colliding || all_goals || oob || timeout;      // ship_env.py:115-134

        if (leader) st_goal += goal_reached ? 1.f : 0.f;
        const bool do_reset = done && p.auto_reset;
        if (done) {
            if (leader) {
                st_episodes += 1.f; st_return += r.ret; st_length += (float)r.steps;
                st_coll += colliding ? 1.f : 0.f; st_oob += oob ? 1.f : 0.f;
                st_timeout += timeout ? 1.f : 0.f; st_allgoals += all_goals ? 1.f : 0.f;
            }
            if (do_reset) {
                const int ep = r.episode + 1;
                reset_env(p, r, pick_scenario(p, gid, ep), ep);
                load_scen_consts(p, r.scen, sc);
                c = 1.f; s = 0.f;
                hminx = p.ship_aabb[0]; hminy = p.ship_aabb[1]; hmaxx = p.ship_aabb[2]; hmaxy = p.ship_aabb[3];
                closest_goal(r, gx, gy);
                goals_dirty = true;
            }
        }

        // ---- outputs.  The leader completes the newest frame in the resident tile (lidar slots are already there);
        // obs rows of the warp's envs are contiguous in global memory, so the tile is copied out with fully
        // coalesced 128-bit streaming stores.
        __syncwarp();                                           // ray hits of the other lanes are visible
        if (gl == 0) {
            if (do_reset) {                                     // ship_env.py:180-184: [-1 x 16 | reset frame], vals = -1
                const float4 neg = make_float4(-1.f, -1.f, -1.f, -1.f);
                if (HIST == 2) { row4[0] = neg; row4[1] = neg; row4[2] = neg; row4[3] = neg; }
                row4[OBS4 - 3] = make_float4(gx, gy, -1.f, -1.f);
                row4[OBS4 - 2] = neg;
                row4[OBS4 - 1] = neg;
            } else {
                reinterpret_cast<float2 *>(row4 + OBS4 - 3)[0] = make_float2(gx, gy);
            }
            row4[OBS4 - 4] = make_float4(r.x, r.y, (float)r.rudder, r.th);
        }
        __syncwarp();
        if (p.obs) {
            float4 *o = p.obs + ((size_t)k * p.N + warp_env0) * OBS4;
            const int n_rows = min(EPW, p.N - warp_env0);
#pragma unroll
            for (int f = lane; f < EPW * OBS4; f += 32) {
                const int rr = f / OBS4, cc = f - rr * OBS4;
                if (rr < n_rows) __stcs(o + f, tile[rr * ROW4 + cc]);
            }
        }
        if (leader) {
            const size_t row = (size_t)k * p.N + e;
            if (p.reward) p.reward[row] = reward;
            if (p.done) p.done[row] = done ? 1 : 0;
        }
        __syncwarp();                                           // copy-out done before the next step rewrites the tile
    }
    if (leader) {
        const float4 l1 = row4[OBS4 - 3], l2 = row4[OBS4 - 2], l3 = row4[OBS4 - 1];
        r.lid[0] = l1.z; r.lid[1] = l1.w; r.lid[2] = l2.x; r.lid[3] = l2.y; r.lid[4] = l2.z; r.lid[5] = l2.w;
        r.lid[6] = l3.x; r.lid[7] = l3.y; r.lid[8] = l3.z; r.lid[9] = l3.w;
        store_env(p, e, r, goals_dirty);
    }

    // episode statistics: warp shuffle reduction, then one red.add per non-zero value per warp into a slot row
    float v[8] = {st_episodes, st_return, st_length, st_goal, st_coll, st_oob, st_timeout, st_allgoals};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(kFull, v[i], off);
    }
    if (lane == 0 && p.stats) {
        double *srow = p.stats + (size_t)(blockIdx.x % kStatSlots) * kStatLen;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (v[i] != 0.f) atomicAdd(srow + i, (double)v[i]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// reset kernel: ShipEnv.reset for the masked envs (ship_env.py:171-184)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reset_kernel(const __grid_constant__ StepParams p, const uint8_t *mask,
                                                    const int *scenario, int first, float4 *obs)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.N) return;
    if (mask && !mask[e]) return;
    EnvRegs r;
    load_env(p, e, r);
    const int ep = first ? 0 : r.episode + 1;
    const int scen = scenario ? scenario[e] : pick_scenario(p, p.env_id_offset + e, ep);
    reset_env(p, r, scen, ep);
    store_env(p, e, r, true);
    if (obs) {
        float gx, gy;
        closest_goal(r, gx, gy);
        float4 *o = obs + (size_t)e * (4 * p.history);
        const float4 neg = make_float4(-1.f, -1.f, -1.f, -1.f);
        if (p.history == 2) { o[0] = neg; o[1] = neg; o[2] = neg; o[3] = neg; o += 4; }
        o[0] = make_float4(r.x, r.y, 0.f, 0.f);
        o[1] = make_float4(gx, gy, -1.f, -1.f);
        o[2] = neg;
        o[3] = neg;
    }
}

// stats slots -> out[kStatLen]; one warp per statistic column
__global__ void stats_reduce_kernel(double *slots, double *out, int clear)
{
    const int col = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int s = lane; s < kStatSlots; s += 32) {
        acc += slots[s * kStatLen + col];
        if (clear) slots[s * kStatLen + col] = 0.0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
    if (lane == 0) out[col] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// launch wrappers (called from the C ABI)
// ------------------------------------------------------------------------------------------------------------
template <int G>
static cudaError_t launch_g(const StepParams &p, cudaStream_t stream, LaunchShape *shape)
{
    const int envs_per_cta = kThreads / G;
    const int blocks = (p.N + envs_per_cta - 1) / envs_per_cta;
    if (shape) { shape->lanes_per_env = G; shape->threads = kThreads; shape->blocks = blocks; }
    if (p.history == 2) step_kernel<G, 2><<<blocks, kThreads, 0, stream>>>(p);
    else step_kernel<G, 1><<<blocks, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_step(const StepParams &p, int lanes_per_env, cudaStream_t stream, LaunchShape *shape)
{
    switch (lanes_per_env) {
        case 1: return launch_g<1>(p, stream, shape);
        case 2: return launch_g<2>(p, stream, shape);
        case 4: return launch_g<4>(p, stream, shape);
        case 8: return launch_g<8>(p, stream, shape);
        case 16: return launch_g<16>(p, stream, shape);
        case 32: return launch_g<32>(p, stream, shape);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_reset(const StepParams &p, const uint8_t *mask, const int *scenario, int first, float4 *obs,
                         cudaStream_t stream)
{
    const int threads = 256;
    reset_kernel<<<(p.N + threads - 1) / threads, threads, 0, stream>>>(p, mask, scenario, first, obs);
    return cudaGetLastError();
}

cudaError_t launch_stats_reduce(double *slots, double *out, int clear, cudaStream_t stream)
{
    stats_reduce_kernel<<<1, 32 * kStatLen, 0, stream>>>(slots, out, clear);
    return cudaGetLastError();
}

}  // namespace shipsim
