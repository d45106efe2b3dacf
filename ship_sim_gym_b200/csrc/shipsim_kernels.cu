// shipsim_kernels.cu -- the fused ShipEnv step kernel (K env-steps per launch), reset and stats kernels.
// sm_100a only.  See shipsim_device.cuh for the data layout, the execution model (G lanes per env, warp-wide
// cooperative geometry) and the reference lines each piece restates.
#include "shipsim_device.cuh"
#include "shipsim_launch.h"

namespace shipsim {

constexpr int kObsRow4 = 9;     // obs staging row stride in float4 (144 B): conflict-free 128-bit shared stores

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

// ------------------------------------------------------------------------------------------------------------
// G lanes per env, 32/G envs per warp.  State is loaded once, lives in registers for K steps, stored once.
// ------------------------------------------------------------------------------------------------------------
template <int G, int HIST>
__global__ void __launch_bounds__(kThreads) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int EPW = 32 / G;                 // envs per warp
    constexpr int OBS4 = 4 * HIST;              // float4 per obs row
    constexpr unsigned GMASK = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    __shared__ float4 s_obs[(kThreads / 32) * EPW * kObsRow4];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int gshift = grp * G;
    const int warp_env0 = (blockIdx.x * (kThreads / 32) + warp) * EPW;
    const int e = warp_env0 + grp;
    const bool valid = e < p.N;
    const bool leader = valid && gl == 0;
    float4 *tile = s_obs + warp * EPW * kObsRow4;

    // lanes <-> (edge slot, ray) in the cooperative ray pass: lane = slot*10 + ray
    const int my_slot = lane / kBeams;
    const int my_ray = lane - my_slot * kBeams;
    const float my_rc = p.ray_c[my_ray], my_rs = p.ray_s[my_ray];
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

#pragma unroll 1
    for (int k = 0; k < p.K; ++k) {
        const int a = load_action(p, k, valid ? e : p.N - 1, gid);
        // previous frame = newest frame of the last step / reset (SURVEY.md App. A note N2)
        float4 P0 = make_float4(r.x, r.y, (float)r.rudder, r.th);
        float4 P1 = make_float4(gx, gy, r.lid[0], r.lid[1]);
        float4 P2 = make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]);
        float4 P3 = make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]);
        const float4 *rec = p.bank + (size_t)r.scen * p.scen_stride4;
        const float4 *E0 = rec + kBankHeader4, *E1 = E0 + p.maxv;

        // ---- ShipGame.handle_discrete_action (game.py:140-153); Ship.move_forward / rotate (models.py:129-146)
        float dvx = 0.f, dvy = 0.f, dw = 0.f;
        if (a == 0) { dvx = -p.acc_dt * s; dvy = p.acc_dt * c; dw = -p.ang_dt * (float)r.rudder; }
        else if (a == 1) r.rudder = max(r.rudder - 5, -10);
        else if (a == 2) r.rudder = min(r.rudder + 5, 10);

        // ---- LiDAR.query (models.py:39-76) at the PRE-integration pose (game.py:193 precedes :194)
        // ray origin = body origin + half the extents of the hull's cached AABB (models.py:51-53)
        float hx, hy;
        {
            float minx = 0.f, maxx = 0.f, miny = 0.f, maxy = 0.f;      // hull vertex 0 is the body origin
#pragma unroll
            for (int j = 1; j < kShipVerts; ++j) {
                const float wx = p.ship_lx[j] * c - p.ship_ly[j] * s;
                const float wy = p.ship_lx[j] * s + p.ship_ly[j] * c;
                minx = fminf(minx, wx); maxx = fmaxf(maxx, wx); miny = fminf(miny, wy); maxy = fmaxf(maxy, wy);
            }
            hx = 0.5f * (maxx - minx); hy = 0.5f * (maxy - miny);
        }
        bool reach0, reach1;
        {
            // box of the fan (origin + the 10 ray ends) against the bank boxes: a bank the fan cannot touch is skipped
            float fx0 = 0.f, fx1 = 0.f, fy0 = 0.f, fy1 = 0.f;
#pragma unroll
            for (int j = 0; j < kBeams; ++j) {
                const float dx = c * p.ray_c[j] - s * p.ray_s[j], dy = s * p.ray_c[j] + c * p.ray_s[j];
                fx0 = fminf(fx0, dx); fx1 = fmaxf(fx1, dx); fy0 = fminf(fy0, dy); fy1 = fmaxf(fy1, dy);
            }
            const float ox = r.x + hx, oy = r.y + hy;
            const float lx = ox + L * fx0, ux = ox + L * fx1, ly = oy + L * fy0, uy = oy + L * fy1;
            reach0 = valid && !(ux < sc.bb0.x || lx > sc.bb0.z || uy < sc.bb0.y || ly > sc.bb0.w);
            reach1 = valid && !(ux < sc.bb1.x || lx > sc.bb1.z || uy < sc.bb1.y || ly > sc.bb1.w);
        }
        // pass 1 (cpShapePointQuery + plane culling): which planes lie within reach in front of the origin, and is
        // the origin inside the bank?  Edges are strided over the G lanes of the group.
        unsigned live0 = 0u, live1 = 0u, out0 = 0u, out1 = 0u;
        if (G == 1) {
            if (reach0)
                for (int i = 0; i < sc.n0; ++i) {
                    const float4 ed = __ldg(E0 + i);
                    const float d = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);     // n.(origin - v_i)
                    out0 |= (d > 0.f) ? 1u : 0u;
                    live0 |= (d >= 0.f && d <= L) ? (1u << i) : 0u;
                }
            if (reach1)
                for (int i = 0; i < sc.n1; ++i) {
                    const float4 ed = __ldg(E1 + i);
                    const float d = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy);
                    out1 |= (d > 0.f) ? 1u : 0u;
                    live1 |= (d >= 0.f && d <= L) ? (1u << i) : 0u;
                }
        } else {
            for (int it = 0; it * G < p.maxv; ++it) {                  // warp-uniform trip count
                const int i = it * G + gl;
                const bool a0 = reach0 && i < sc.n0, a1 = reach1 && i < sc.n1;
                float d0 = -1.f, d1 = -1.f;
                if (a0) { const float4 ed = __ldg(E0 + i); d0 = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy); }
                if (a1) { const float4 ed = __ldg(E1 + i); d1 = ed.x * ((r.x - ed.z) + hx) + ed.y * ((r.y - ed.w) + hy); }
                const unsigned bl0 = __ballot_sync(kFull, a0 && d0 >= 0.f && d0 <= L), bo0 = __ballot_sync(kFull, a0 && d0 > 0.f);
                const unsigned bl1 = __ballot_sync(kFull, a1 && d1 >= 0.f && d1 <= L), bo1 = __ballot_sync(kFull, a1 && d1 > 0.f);
                live0 |= ((bl0 >> gshift) & GMASK) << (it * G); out0 |= (bo0 >> gshift) & GMASK;
                live1 |= ((bl1 >> gshift) & GMASK) << (it * G); out1 |= (bo1 >> gshift) & GMASK;
            }
        }
        const bool in0 = reach0 && out0 == 0u, in1 = reach1 && out1 == 0u;   // origin inside the bank polygon

        // pass 2, cooperative: the whole warp serves one env at a time.  Lane = slot*10 + ray: up to three live
        // edges x ten rays per round.  cpPolyShapeSegmentQuery: later edges overwrite earlier ones; LiDAR.query:
        // the first bank (list order) that reports a hit wins, misses keep the old reading (sticky vals).
        {
            unsigned needy = __ballot_sync(kFull, gl == 0 && (in0 || in1 || live0 != 0u || live1 != 0u));
            while (needy) {
                const int src = __ffs(needy) - 1;
                needy &= needy - 1u;
                const float bx = __shfl_sync(kFull, r.x, src), by = __shfl_sync(kFull, r.y, src);
                const float bhx = __shfl_sync(kFull, hx, src), bhy = __shfl_sync(kFull, hy, src);
                const float bc = __shfl_sync(kFull, c, src), bs = __shfl_sync(kFull, s, src);
                const int bscen = __shfl_sync(kFull, r.scen, src);
                const int bflags = __shfl_sync(kFull, sc.n0 | (sc.n1 << 8) | ((int)in0 << 16) | ((int)in1 << 17), src);
                const unsigned blive0 = __shfl_sync(kFull, live0, src), blive1 = __shfl_sync(kFull, live1, src);
                const float4 *bE = p.bank + (size_t)bscen * p.scen_stride4 + kBankHeader4;
                const float dirx = bc * my_rc - bs * my_rs, diry = bs * my_rc + bc * my_rs;   // cos/sin(angle + a_ray)
                unsigned pend = (1u << kBeams) - 1u;
                bool hit = false;
                float val = 0.f;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const bool inside = (bflags >> (16 + b)) & 1;
                    unsigned lv = b == 0 ? blive0 : blive1;
                    const int nb = (bflags >> (8 * b)) & 0xff;
                    const float4 *E = bE + b * p.maxv;
                    if (inside) {            // start point inside the shape: alpha = 0, point stays at the ray end
                        if (lane < kBeams && (pend >> lane & 1u)) { hit = true; val = L; }
                        pend = 0u;
                    } else {
                        bool hit_b = false;
                        float val_b = 0.f;
                        while (lv != 0u && pend != 0u) {
                            const unsigned pos = __fns(lv, 0, my_slot + 1);
                            int ok = 0;
                            float dist = 0.f;
                            if (pos < 32u && lane < 3 * kBeams && (pend >> my_ray & 1u)) {
                                const int i = (int)pos;
                                const float4 ed = __ldg(E + i);
                                const float4 ep = __ldg(E + (i == 0 ? nb - 1 : i - 1));
                                // The hit distance is d / (-n.dir): an error of d is amplified by 1/cos(incidence).
                                // The stored fp32 normal is good to ~6e-8 rad, i.e. 6e-5 of d over a 1000-unit
                                // edge, so for live edges the plane is rebuilt in double from the two fp32
                                // vertices (what the reference's double planes are made of).  FP64 runs at half
                                // the FP32 rate on B200 and only a few lanes-steps get here.
                                const double exd = (double)ed.z - (double)ep.z, eyd = (double)ed.w - (double)ep.w;
                                const double len2 = exd * exd + eyd * eyd;
                                const double inv = rsqrt(len2);
                                const double nxd = eyd * inv, nyd = -exd * inv;
                                const double qxd = ((double)bx - (double)ed.z) + (double)bhx;
                                const double qyd = ((double)by - (double)ed.w) + (double)bhy;
                                const float d = (float)(nxd * qxd + nyd * qyd);
                                const float ta = (float)(nxd * qyd - nyd * qxd);       // cross(n, origin - v_i)
                                const float tmin = -(float)(len2 * inv);               // cross(n, v_{i-1} - v_i) = -|edge|
                                const float enx = (float)nxd, eny = (float)nyd;
                                const float denom = -L * (enx * dirx + eny * diry);    // an - bn
                                float t;
                                if (denom > 0.f) t = __fdividef(d, denom); else t = (d == 0.f) ? 0.f : 2.f;   // d / max(an-bn, DBL_MIN)
                                const float tang = ta + t * L * (enx * diry - eny * dirx);                  // cross(n, hit - v_i)
                                ok = (d >= 0.f) && (t <= 1.f) && (tang >= tmin) && (tang <= 0.f);
                                dist = t * L;
                            }
                            const int ok1 = __shfl_down_sync(kFull, ok, kBeams), ok2 = __shfl_down_sync(kFull, ok, 2 * kBeams);
                            const float d1 = __shfl_down_sync(kFull, dist, kBeams), d2 = __shfl_down_sync(kFull, dist, 2 * kBeams);
                            if (lane < kBeams) {
                                if (ok2) { hit_b = true; val_b = d2; }
                                else if (ok1) { hit_b = true; val_b = d1; }
                                else if (ok) { hit_b = true; val_b = dist; }
                            }
                            lv &= lv - 1u; lv &= lv - 1u; lv &= lv - 1u;      // the three lowest live edges are done
                        }
                        const unsigned hm = __ballot_sync(kFull, lane < kBeams && hit_b);
                        if (lane < kBeams && hit_b) { hit = true; val = val_b; }
                        pend &= ~hm;
                    }
                }
                const unsigned hitmask = __ballot_sync(kFull, lane < kBeams && hit);
                if (hitmask) {
                    const bool mine = (lane / G) == (src / G);
#pragma unroll
                    for (int j = 0; j < kBeams; ++j) {
                        const float v = __shfl_sync(kFull, val, j);
                        if (mine && (hitmask >> j & 1u)) r.lid[j] = v;
                    }
                }
            }
        }

        // ---- cpSpaceStep: positions first (cpBodyUpdatePosition)
        r.x += r.vx * p.dt;
        r.y += r.vy * p.dt;
        r.th += r.w * p.dt;
        sincos_fast(r.th, s, c);

        // ---- overlap tests at the new pose -> begin callbacks collide_ship / collide_goal (game.py:232-257)
        bool ov0, ov1;
        {
            float sminx = 0.f, sminy = 0.f, smaxx = 0.f, smaxy = 0.f;
#pragma unroll
            for (int j = 1; j < kShipVerts; ++j) {
                const float wx = p.ship_lx[j] * c - p.ship_ly[j] * s;
                const float wy = p.ship_lx[j] * s + p.ship_ly[j] * c;
                sminx = fminf(sminx, wx); smaxx = fmaxf(smaxx, wx); sminy = fminf(sminy, wy); smaxy = fmaxf(smaxy, wy);
            }
            sminx += r.x; smaxx += r.x; sminy += r.y; smaxy += r.y;
            // cpBBIntersects (inclusive) pre-filter of the narrow phase
            ov0 = valid && !(sminx > sc.bb0.z || smaxx < sc.bb0.x || sminy > sc.bb0.w || smaxy < sc.bb0.y);
            ov1 = valid && !(sminx > sc.bb1.z || smaxx < sc.bb1.x || sminy > sc.bb1.w || smaxy < sc.bb1.y);
        }
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
        {
            // goals: cheap cull in the body frame for all five, exact distance test only for the survivors
            float qx[kGoals], qy[kGoals];
            unsigned cand = 0u;
#pragma unroll
            for (int g = 0; g < kGoals; ++g) {
                const float ux = r.g[2 * g] - r.x, uy = r.g[2 * g + 1] - r.y;
                qx[g] = ux * c + uy * s; qy[g] = -ux * s + uy * c;
                if (valid && ((r.alive >> g) & 1) && !goal_culled(p, qx[g], qy[g])) cand |= 1u << g;
            }
            while (cand) {
                const int g = __ffs(cand) - 1;
                cand &= cand - 1u;
                float ax = qx[0], ay = qy[0];
#pragma unroll
                for (int j = 1; j < kGoals; ++j) if (g == j) { ax = qx[j]; ay = qy[j]; }
                if (goal_touches_ship(p, ax, ay)) { goal_reached = true; r.alive &= ~(1 << g); }
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
        closest_goal(r, gx, gy);
        const bool all_goals = (r.alive == 0);
        const bool timeout = (r.steps >= p.max_steps);
        const bool done = colliding || all_goals || oob || timeout;      // ship_env.py:115-134

        if (leader) st_goal += goal_reached ? 1.f : 0.f;
        if (done) {
            if (leader) {
                st_episodes += 1.f; st_return += r.ret; st_length += (float)r.steps;
                st_coll += colliding ? 1.f : 0.f; st_oob += oob ? 1.f : 0.f;
                st_timeout += timeout ? 1.f : 0.f; st_allgoals += all_goals ? 1.f : 0.f;
            }
            if (p.auto_reset) {
                const int ep = r.episode + 1;
                reset_env(p, r, pick_scenario(p, gid, ep), ep);
                load_scen_consts(p, r.scen, sc);
                c = 1.f; s = 0.f;
                closest_goal(r, gx, gy);
                goals_dirty = true;
                P0 = P1 = P2 = P3 = make_float4(-1.f, -1.f, -1.f, -1.f);   // ship_env.py:180-181
            }
        }

        // ---- outputs.  obs rows of the warp's envs are contiguous in global memory: stage them in shared memory
        // (padded rows, conflict-free) and write them back with fully coalesced 128-bit stores.
        const size_t row = (size_t)k * p.N + e;
        if (p.obs) {
            __syncwarp();
            if (gl == 0) {
                float4 *t = tile + grp * kObsRow4;
                if (HIST == 2) { t[0] = P0; t[1] = P1; t[2] = P2; t[3] = P3; t += 4; }
                t[0] = make_float4(r.x, r.y, (float)r.rudder, r.th);
                t[1] = make_float4(gx, gy, r.lid[0], r.lid[1]);
                t[2] = make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]);
                t[3] = make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]);
            }
            __syncwarp();
            float4 *o = p.obs + ((size_t)k * p.N + warp_env0) * OBS4;
            const int n_rows = min(EPW, p.N - warp_env0);
#pragma unroll
            for (int f = lane; f < EPW * OBS4; f += 32) {
                const int rr = f / OBS4, cc = f - rr * OBS4;
                if (rr < n_rows) __stcs(o + f, tile[rr * kObsRow4 + cc]);
            }
        }
        if (leader) {
            if (p.reward) p.reward[row] = reward;
            if (p.done) p.done[row] = done ? 1 : 0;
        }
    }
    if (leader) store_env(p, e, r, goals_dirty);

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
