// shipsim_kernels.cu -- the fused ShipEnv step kernel (K env-steps per launch), reset and stats kernels.
// sm_100a only.  See shipsim_device.cuh for the data layout, the execution model (G lanes per env, warp-wide
// cooperative geometry) and the reference lines each piece restates.
#include "shipsim_device.cuh"
#include "shipsim_launch.h"

#ifndef SHIPSIM_MIN_BLOCKS
#define SHIPSIM_MIN_BLOCKS 4
#endif

namespace shipsim {

// actions[k][e]: `ap` walks down this env's column (stride = one row of the action tensor, in bytes)
__device__ __forceinline__ int load_action(const StepParams &p, const char *ap, int k, long long gid)
{
    switch (p.action_dtype) {
        case 0: return __ldg(reinterpret_cast<const int *>(ap));
        case 1: return (int)__ldg(reinterpret_cast<const long long *>(ap));
        case 2: return (int)__ldg(reinterpret_cast<const unsigned char *>(ap));
        default: return random_action(p, gid, p.step0 + (unsigned)k);
    }
}

// One bank-normal axis of the separating-axis test: does plane `ed` of the bank have the whole ship (rotated hull
// rx/ry about the body origin bx/by) strictly in front of it?  Explicit fma/mul so that every call site rounds alike.
__device__ __forceinline__ bool bank_axis_separates(const float4 ed, const float (&rx)[kShipVerts], const float (&ry)[kShipVerts],
                                                    float bx, float by)
{
    float m = fmaf(ed.x, rx[0], __fmul_rn(ed.y, ry[0]));
#pragma unroll
    for (int j = 1; j < kShipVerts; ++j) m = fminf(m, fmaf(ed.x, rx[j], __fmul_rn(ed.y, ry[j])));
    const float base = fmaf(ed.x, bx - ed.z, __fmul_rn(ed.y, by - ed.w));
    return base + m > 0.f;
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

// Reach-grid cell of the lidar origin (ox, oy): which bank edges a ray starting there can touch at all.
__device__ __forceinline__ uint4 load_cell(const StepParams &p, int scen, float ox, float oy)
{
    int ix = __float2int_rd((ox - p.gridp.x0) * p.gridp.inv_cx);
    int iy = __float2int_rd((oy - p.gridp.y0) * p.gridp.inv_cy);
    ix = min(max(ix, 0), kGridN - 1);
    iy = min(max(iy, 0), kGridN - 1);
    return __ldg(p.grid + ((size_t)scen * kGridN + iy) * kGridN + ix);
}

constexpr int kMaxCand = 4;          // candidate planes per env the shared-memory scratch holds (more -> serial path)

// A candidate plane seen from the ray origin (ox, oy) = (x + hx, y + hy), evaluated in double from the plane the
// reference's cpSplittingPlane holds: d = n.(o - v_i), ta = cross(n, o - v_i); n is returned scaled by -L so that
// the per-ray part needs no further multiplies.
struct PlaneEval { float d, ta, nxl, nyl, len; };

__device__ __forceinline__ PlaneEval eval_plane(const EdgeD *E, double xd, double yd, double hxd, double hyd, float L)
{
    const double2 nd = __ldg(reinterpret_cast<const double2 *>(E));
    const float4 ev = __ldg(reinterpret_cast<const float4 *>(E) + 1);
    const double qx = (xd - (double)ev.x) + hxd, qy = (yd - (double)ev.y) + hyd;     // origin - v_i
    PlaneEval o;
    o.d = (float)(nd.x * qx + nd.y * qy);
    o.ta = (float)(nd.x * qy - nd.y * qx);
    o.nxl = -L * (float)nd.x;
    o.nyl = -L * (float)nd.y;
    o.len = ev.z;
    return o;
}

// One ray against one candidate plane (cpPolyShapeSegmentQuery's loop body).  t = d / max(an - bn, DBL_MIN) and the
// test t <= 1 is evaluated as d <= an - bn; `hit` <=> the reference accepts this edge, `val` = |hit - origin|.
__device__ __forceinline__ bool ray_vs_plane(const PlaneEval &e, float dirx, float diry, float L, float &val)
{
    const float denom = e.nxl * dirx + e.nyl * diry;                      // an - bn
    const float cr = e.nxl * diry - e.nyl * dirx;                         // -L * cross(n, dir)
    const bool pos = denom > 0.f;
    const float t = pos ? __fdividef(e.d, denom) : 0.f;
    const float tang = e.ta - t * cr;                                     // cross(n, hit - v_i)
    val = t * L;
    return e.d >= 0.f && (pos ? e.d <= denom : e.d == 0.f) && tang >= -e.len && tang <= 0.f;
}

// Serial LiDAR.query of one env by one lane: only for cells with more than kMaxCand candidate planes.
__device__ __noinline__ void ray_query_serial(const StepParams &p, float x, float y, float hx, float hy, float c, float s, int scen,
                                              unsigned m0, unsigned m1, unsigned flags, float *lid)
{
    const float L = p.lidar_len;
    const EdgeD *E = p.edges_d + (size_t)scen * (2 * kMaxHull);
    const double xd = (double)x, yd = (double)y, hxd = (double)hx, hyd = (double)hy;
    unsigned pend = (1u << kBeams) - 1u;
    for (int b = 0; b < 2; ++b) {
        const unsigned mb = b ? m1 : m0;
        bool out = false;
        unsigned hitm = 0u;
        float v[kBeams];
        for (unsigned m = mb; m; m &= m - 1u) {
            const PlaneEval pe = eval_plane(E + b * kMaxHull + (__ffs(m) - 1), xd, yd, hxd, hyd, L);
            out = out || (pe.d > 0.f);
#pragma unroll
            for (int j = 0; j < kBeams; ++j) {
                const float dirx = c * p.ray_c[j] - s * p.ray_s[j], diry = s * p.ray_c[j] + c * p.ray_s[j];
                float val;
                if (ray_vs_plane(pe, dirx, diry, L, val)) { v[j] = val; hitm |= 1u << j; }
            }
        }
        const bool inside = ((flags >> b) & 1u) && !out;
#pragma unroll
        for (int j = 0; j < kBeams; ++j)
            if ((pend >> j) & 1u) {
                if (inside) lid[j] = L;
                else if ((hitm >> j) & 1u) lid[j] = v[j];
            }
        pend &= inside ? 0u : ~hitm;
    }
}

// ------------------------------------------------------------------------------------------------------------
// G lanes per env, 32/G envs per warp.  The lanes of a group hold identical copies of the env's scalar state
// (loaded once, in registers for K steps, stored once).  The two most recent observation frames of every env --
// including the sticky lidar readings, which are state -- live in a padded shared-memory tile from which each
// step's obs rows are copied out, fully coalesced.  The irregular geometry is done by the WHOLE WARP in passes:
//   ray pass: 3 needy envs at a time, one lane per (env, ray); the reach grid supplies the candidate edges
//   SAT pass: 32/lps needy envs at a time, one lane per (env, bank edge), lps = 8, 16 or 32 >= max hull size
// so that control flow stays warp-uniform no matter how few envs of a warp are near a bank.
// ------------------------------------------------------------------------------------------------------------
template <int G, int HIST>
__global__ void __launch_bounds__(kThreads, SHIPSIM_MIN_BLOCKS) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int EPW = 32 / G;                 // envs per warp
    constexpr int OBS4 = 4 * HIST;              // float4 per obs row
    constexpr int ROW4 = OBS4 + 1;              // padded tile row (odd float4 stride: conflict-free 128-bit accesses)
    constexpr int CF = 16 * (HIST - 1);         // float offset of the newest frame inside a row
    constexpr int SCR4 = 1 + 2 * kMaxCand;      // scratch row: header + two float4 per candidate plane (odd stride)
    __shared__ float4 s_tile[(kThreads / 32) * EPW * ROW4];
    __shared__ float4 s_scr[(kThreads / 32) * EPW * SCR4];

    __shared__ float s_ray[2 * 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int warp_env0 = (blockIdx.x * (kThreads / 32) + warp) * EPW;
    const int e = warp_env0 + grp;
    const bool valid = e < p.N;
    const bool leader = valid && gl == 0;
    float4 *tile = s_tile + warp * EPW * ROW4;
    float4 *row4 = tile + grp * ROW4;                               // this env's [older frame | newest frame]
    float *tile_f = reinterpret_cast<float *>(tile);
    float4 *scr = s_scr + warp * EPW * SCR4;
    const float L = p.lidar_len;

    // lane roles in the cooperative passes
    const int rslot = lane / kBeams;                                // ray pass: env slot 0..2 (lanes 30, 31 idle)
    if (threadIdx.x < 32) {                                         // per-lane ray direction table (body frame)
        const int j = threadIdx.x % kBeams;
        s_ray[threadIdx.x] = p.ray_c[j];
        s_ray[32 + threadIdx.x] = p.ray_s[j];
    }
    const int lps_sh = p.maxv <= 8 ? 3 : (p.maxv <= 16 ? 4 : 5);    // SAT pass: log2(lanes per env slot)
    const int lps = 1 << lps_sh, nslots = 32 >> lps_sh;
    const int sslot = lane >> lps_sh, sel = lane & (lps - 1);
    const unsigned slotmask = lps == 32 ? kFull : (((1u << lps) - 1u) << (sslot * lps));

    const int cp_row0 = lane / OBS4, cp_src0 = cp_row0 * ROW4 + lane % OBS4;      // obs copy-out: this lane's first float4
    const int n_rows = min(EPW, p.N - warp_env0);

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
    float hminx, hmaxx, hminy, hmaxy;
    hull_extents(p, c, s, hminx, hmaxx, hminy, hmaxy);
    // ray origin = body origin + half the extents of the hull's cached AABB (models.py:51-53)
    float hx = 0.5f * (hmaxx - hminx), hy = 0.5f * (hmaxy - hminy);
    uint4 cell = load_cell(p, r.scen, r.x + hx, r.y + hy);
    int axis0 = 0, axis1 = 0;                    // last separating bank edge per bank (temporal coherence; never stored)
    if (gl == 0) {                               // newest frame of the resident tile = frame of the current state
        row4[OBS4 - 4] = make_float4(r.x, r.y, (float)r.rudder, r.th);
        row4[OBS4 - 3] = make_float4(gx, gy, r.lid[0], r.lid[1]);
        row4[OBS4 - 2] = make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]);
        row4[OBS4 - 1] = make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]);
    }
    const size_t act_esize = p.action_dtype == 1 ? 8 : (p.action_dtype == 2 ? 1 : 4);
    const size_t act_stride = (size_t)p.N * act_esize;
    const char *ap = reinterpret_cast<const char *>(p.actions) + (size_t)(valid ? e : p.N - 1) * act_esize;
    int a_next = load_action(p, ap, 0, gid);
    __syncthreads();                             // s_ray visible; also orders the tile initialisation

#pragma unroll 1
    for (int k = 0; k < p.K; ++k) {
        const int a = a_next;
        ap += act_stride;
        if (k + 1 < p.K) a_next = load_action(p, ap, k + 1, gid);                     // prefetch: off the critical path
        // previous frame <- newest frame of the last step / reset (SURVEY.md App. A note N2); lidar stays in place
        if (HIST == 2 && gl == 0) { row4[0] = row4[4]; row4[1] = row4[5]; row4[2] = row4[6]; row4[3] = row4[7]; }

        // ---- ShipGame.handle_discrete_action (game.py:140-153); Ship.move_forward / rotate (models.py:129-146)
        float dvx = 0.f, dvy = 0.f, dw = 0.f;
        if (a == 0) { dvx = -p.acc_dt * s; dvy = p.acc_dt * c; dw = -p.ang_dt * (float)r.rudder; }
        else if (a == 1) r.rudder = max(r.rudder - 5, -10);
        else if (a == 2) r.rudder = min(r.rudder + 5, 10);

        // ---- LiDAR.query (models.py:39-76) at the PRE-integration pose (game.py:193 precedes :194)
        // phase 1, per env (owner lane): evaluate the candidate planes the reach grid names, in double, into the scratch row
        const bool needy = leader && ((cell.x | cell.y | (cell.z & 3u)) != 0u);
        const bool big = needy && (__popc(cell.x) + __popc(cell.y) > kMaxCand);
        if (HIST == 2) __syncwarp();            // the frame copy has read the old readings before any lane overwrites them
        if (needy) {
            if (big) {
                ray_query_serial(p, r.x, r.y, hx, hy, c, s, r.scen, cell.x, cell.y, cell.z, tile_f + grp * (ROW4 * 4) + CF + 6);
            } else {
                float4 *row = scr + grp * SCR4;
                const EdgeD *E = p.edges_d + (size_t)r.scen * (2 * kMaxHull);
                const double xd = (double)r.x, yd = (double)r.y, hxd = (double)hx, hyd = (double)hy;
                unsigned m0 = cell.x, m1 = cell.y;
                int n = 0;
                bool out0 = false, out1 = false;
                while (m0 | m1) {
                    int idx;
                    if (m0) { idx = __ffs(m0) - 1; m0 &= m0 - 1u; } else { idx = kMaxHull + __ffs(m1) - 1; m1 &= m1 - 1u; }
                    const PlaneEval pe = eval_plane(E + idx, xd, yd, hxd, hyd, L);
                    if (idx < kMaxHull) out0 = out0 || (pe.d > 0.f); else out1 = out1 || (pe.d > 0.f);
                    row[1 + 2 * n] = make_float4(pe.d, pe.ta, pe.nxl, pe.nyl);
                    row[2 + 2 * n] = make_float4(pe.len, idx < kMaxHull ? 0.f : 1.f, 0.f, 0.f);
                    ++n;
                }
                // cpShapeSegmentQuery: start point inside the shape => alpha = 0 and `point` stays at the ray end
                const unsigned in0 = ((cell.z & 1u) && !out0) ? 1u : 0u, in1 = ((cell.z & 2u) && !out1) ? 1u : 0u;
                row[0] = make_float4(c, s, __int_as_float(n | (int)(in0 << 8) | (int)(in1 << 9)), 0.f);
            }
        }
        // phase 2, whole warp: up to three needy envs per pass, lanes 0-9 / 10-19 / 20-29 = the ten rays of slot 0 / 1 / 2
        unsigned need = __ballot_sync(kFull, needy && !big);
        __syncwarp();
        while (need) {
            const int s0 = __ffs(need) - 1;
            const unsigned n1 = need & (need - 1u);
            const int s1 = __ffs(n1) - 1;                                 // -1 when there is no second env
            const unsigned n2 = n1 & (n1 - 1u);
            const int s2 = __ffs(n2) - 1;
            need = n2 & (n2 - 1u);
            const int src = rslot == 0 ? s0 : (rslot == 1 ? s1 : (rslot == 2 ? s2 : -1));
            if (src >= 0) {
                const int env = src / G;
                const float4 *row = scr + env * SCR4;
                const float4 hdr = row[0];
                const float ray_c = s_ray[lane], ray_s = s_ray[32 + lane];
                const float dirx = hdr.x * ray_c - hdr.y * ray_s, diry = hdr.y * ray_c + hdr.x * ray_s;
                const int hz = __float_as_int(hdr.z);
                const int n = hz & 0xff;
                bool hb0 = false, hb1 = false;
                float v0 = L, v1 = L;
                for (int i = 0; i < n; ++i) {       // cpPolyShapeSegmentQuery: later accepted edges overwrite earlier ones
                    const float4 e0 = row[1 + 2 * i], e1 = row[2 + 2 * i];
                    PlaneEval pe;
                    pe.d = e0.x; pe.ta = e0.y; pe.nxl = e0.z; pe.nyl = e0.w; pe.len = e1.x;
                    float val;
                    const bool ok = ray_vs_plane(pe, dirx, diry, L, val);
                    if (e1.y == 0.f) { if (ok) { v0 = val; hb0 = true; } }
                    else { if (ok) { v1 = val; hb1 = true; } }
                }
                const bool in0 = (hz >> 8) & 1, in1 = (hz >> 9) & 1;
                const bool hit0 = hb0 || in0, hit1 = hb1 || in1;
                // LiDAR.query: the first bank (list order) that reports a hit wins; misses keep the old reading
                // (sticky vals, models.py:71)
                if (hit0 || hit1) tile_f[env * (ROW4 * 4) + CF + 6 + (lane - rslot * kBeams)] = hit0 ? (in0 ? L : v0) : (in1 ? L : v1);
            }
        }

        // ---- cpSpaceStep: positions first (cpBodyUpdatePosition)
        r.x += r.vx * p.dt;
        r.y += r.vy * p.dt;
        r.th += r.w * p.dt;
        sincos_fast(r.th, s, c);
        hull_extents(p, c, s, hminx, hmaxx, hminy, hmaxy);
        hx = 0.5f * (hmaxx - hminx); hy = 0.5f * (hmaxy - hminy);
        cell = load_cell(p, r.scen, r.x + hx, r.y + hy);          // for the NEXT step's lidar; consumed a whole step later

        // ---- overlap tests at the new pose -> begin callbacks collide_ship / collide_goal (game.py:232-257)
        // cpBBIntersects (inclusive) pre-filter of the narrow phase
        const bool ov0 = valid && !(r.x + hminx > sc.bb0.z || r.x + hmaxx < sc.bb0.x || r.y + hminy > sc.bb0.w || r.y + hmaxy < sc.bb0.y);
        const bool ov1 = valid && !(r.x + hminx > sc.bb1.z || r.x + hmaxx < sc.bb1.x || r.y + hminy > sc.bb1.w || r.y + hmaxy < sc.bb1.y);
        bool colliding = false;
        {
            // Separating-axis test.  Contact <=> no separating axis among the edge normals of both convex polygons
            // (touching counts: GJK distance <= 0).  Per lane first: the bank edge that separated last time (or one of
            // its neighbours) almost always still does.  Only envs for which it does not go to the cooperative pass.
            bool ask0 = false, ask1 = false;
            if (ov0 || ov1) {
                float rx[kShipVerts], ry[kShipVerts];
#pragma unroll
                for (int j = 0; j < kShipVerts; ++j) {
                    rx[j] = p.ship_lx[j] * c - p.ship_ly[j] * s;
                    ry[j] = p.ship_lx[j] * s + p.ship_ly[j] * c;
                }
                const float4 *bE = p.bank + (size_t)r.scen * p.scen_stride4 + kBankHeader4;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    if (b ? ov1 : ov0) {
                        const int nb = b ? sc.n1 : sc.n0;
                        int ax = b ? axis1 : axis0;
                        if (ax >= nb) ax = 0;
                        bool sep = false;
#pragma unroll 1
                        for (int tr = 0; tr < 3 && !sep; ++tr) {                // same edge, next edge, previous edge
                            int i = tr == 0 ? ax : (tr == 1 ? ax + 1 : ax - 1);
                            i = i >= nb ? 0 : (i < 0 ? nb - 1 : i);
                            sep = bank_axis_separates(__ldg(bE + b * p.maxv + i), rx, ry, r.x, r.y);
                            if (sep) ax = i;
                        }
                        if (b) { axis1 = ax; ask1 = !sep; } else { axis0 = ax; ask0 = !sep; }
                    }
                }
            }
            // cooperative pass, nslots envs at a time, lanes <-> bank edges
            unsigned needs = __ballot_sync(kFull, gl == 0 && (ask0 || ask1));
            while (needs) {
                int src = -1, myslot = -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < nslots && needs) {
                        const int t = __ffs(needs) - 1;
                        needs &= needs - 1u;
                        if (sslot == q) src = t;
                        if (t / G == grp) myslot = q;
                    }
                }
                const bool act_env = src >= 0;
                const int srcl = act_env ? src : lane;
                const float bx = __shfl_sync(kFull, r.x, srcl), by = __shfl_sync(kFull, r.y, srcl);
                const float bc = __shfl_sync(kFull, c, srcl), bs = __shfl_sync(kFull, s, srcl);
                const int bscen = __shfl_sync(kFull, r.scen, srcl);
                const int bflags = __shfl_sync(kFull, sc.n0 | (sc.n1 << 8) | ((int)ask0 << 16) | ((int)ask1 << 17), srcl);
                const float4 *bE = p.bank + (size_t)bscen * p.scen_stride4 + kBankHeader4;
                float rx[kShipVerts], ry[kShipVerts];
#pragma unroll
                for (int j = 0; j < kShipVerts; ++j) {
                    rx[j] = p.ship_lx[j] * bc - p.ship_ly[j] * bs;
                    ry[j] = p.ship_lx[j] * bs + p.ship_ly[j] * bc;
                }
                bool coll = false;
                unsigned sepbits[2] = {0u, 0u};
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const bool do_b = act_env && ((bflags >> (16 + b)) & 1) && !coll;
                    const int nb = (bflags >> (8 * b)) & 0xff;
                    const bool actl = do_b && sel < nb;
                    float4 ed = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (actl) ed = __ldg(bE + b * p.maxv + sel);
                    sepbits[b] = __ballot_sync(kFull, actl && bank_axis_separates(ed, rx, ry, bx, by));   // a bank edge normal separates
                    bool sep = (sepbits[b] & slotmask) != 0u;
                    if (__ballot_sync(kFull, do_b && !sep)) {                               // rare: try the ship's edge normals
#pragma unroll
                        for (int j = 0; j < kShipVerts; ++j) {
                            const float nx = p.ship_nx[j] * bc - p.ship_ny[j] * bs;
                            const float ny = p.ship_nx[j] * bs + p.ship_ny[j] * bc;
                            const float pr = actl ? nx * (ed.z - bx) + ny * (ed.w - by) : 3.0e38f;
                            const int mn = __reduce_min_sync(slotmask, f2ord(pr));
                            sep = sep || (mn > f2ord(p.ship_off[j]));
                        }
                    }
                    if (do_b && !sep) coll = true;
                }
                const unsigned res = __ballot_sync(kFull, coll);
                if (myslot >= 0) {
                    const int sh = myslot * lps;
                    if ((res >> sh) & 1u) colliding = true;
                    const unsigned sb0 = (sepbits[0] >> sh) & (lps == 32 ? kFull : ((1u << lps) - 1u));
                    const unsigned sb1 = (sepbits[1] >> sh) & (lps == 32 ? kFull : ((1u << lps) - 1u));
                    if (sb0) axis0 = __ffs(sb0) - 1;                                        // remember the separating edges
                    if (sb1) axis1 = __ffs(sb1) - 1;
                }
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
                hx = 0.5f * (hmaxx - hminx); hy = 0.5f * (hmaxy - hminy);
                cell = load_cell(p, r.scen, r.x + hx, r.y + hy);
                axis0 = 0; axis1 = 0;
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
            float4 *o = p.obs + ((size_t)k * p.N + warp_env0) * OBS4 + lane;
            if (EPW * OBS4 >= 32) {
                // lane -> (row lane / OBS4, column lane % OBS4); each further round moves 32 / OBS4 rows down
#pragma unroll
                for (int i = 0; i < EPW * OBS4 / 32; ++i)
                    if (cp_row0 + i * (32 / OBS4) < n_rows) __stcs(o + i * 32, tile[cp_src0 + i * (32 / OBS4) * ROW4]);
            } else if (lane < EPW * OBS4 && cp_row0 < n_rows) {
                __stcs(o, tile[cp_src0]);
            }
        }
        if (leader) {
            const size_t row = (size_t)k * p.N + e;
            if (p.reward) p.reward[row] = reward;
            if (p.done) p.done[row] = done ? 1 : 0;
        }
        {   // the next step's ray pass will want the double planes of these edges: pull them towards L1 now
            const EdgeD *E = p.edges_d + (size_t)r.scen * (2 * kMaxHull);
            if (cell.x) prefetch_l1(E + (__ffs(cell.x) - 1));
            if (cell.y) prefetch_l1(E + kMaxHull + (__ffs(cell.y) - 1));
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
// reach-grid build kernel: one thread per (scenario, cell), double precision.  Runs once per scenario upload.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double point_rect_dist(double px, double py, double x0, double y0, double x1, double y1)
{
    const double dx = fmax(fmax(x0 - px, px - x1), 0.0), dy = fmax(fmax(y0 - py, py - y1), 0.0);
    return sqrt(dx * dx + dy * dy);
}

__device__ __forceinline__ double point_seg_dist(double px, double py, double ax, double ay, double bx, double by)
{
    const double ex = bx - ax, ey = by - ay;
    const double den = ex * ex + ey * ey;
    double t = den > 0.0 ? ((px - ax) * ex + (py - ay) * ey) / den : 0.0;
    t = fmin(fmax(t, 0.0), 1.0);
    const double dx = px - (ax + t * ex), dy = py - (ay + t * ey);
    return sqrt(dx * dx + dy * dy);
}

// distance between the segment a-b and the axis-aligned rectangle [x0,x1] x [y0,y1] (0 when they intersect)
__device__ double seg_rect_dist(double ax, double ay, double bx, double by, double x0, double y0, double x1, double y1)
{
    {   // Liang-Barsky: does any part of the segment lie inside the rectangle?
        double t0 = 0.0, t1 = 1.0;
        const double dx = bx - ax, dy = by - ay;
        const double pp[4] = {-dx, dx, -dy, dy};
        const double qq[4] = {ax - x0, x1 - ax, ay - y0, y1 - ay};
        bool inside = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (pp[i] == 0.0) { if (qq[i] < 0.0) inside = false; }
            else {
                const double t = qq[i] / pp[i];
                if (pp[i] < 0.0) t0 = fmax(t0, t); else t1 = fmin(t1, t);
            }
        }
        if (inside && t0 <= t1) return 0.0;
    }
    double d = fmin(point_rect_dist(ax, ay, x0, y0, x1, y1), point_rect_dist(bx, by, x0, y0, x1, y1));
    d = fmin(d, point_seg_dist(x0, y0, ax, ay, bx, by));
    d = fmin(d, point_seg_dist(x1, y0, ax, ay, bx, by));
    d = fmin(d, point_seg_dist(x0, y1, ax, ay, bx, by));
    d = fmin(d, point_seg_dist(x1, y1, ax, ay, bx, by));
    return d;
}

__global__ void __launch_bounds__(256) build_grid_kernel(const double *hull_xy, const int *hull_n, int n_scen, int maxv_in,
                                                         double gx0, double gy0, double cw, double ch, double reach,
                                                         double touch_margin, uint4 *grid)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_scen * kGridN * kGridN) return;
    const int s = (int)(t / (kGridN * kGridN)), cidx = (int)(t % (kGridN * kGridN));
    const int ix = cidx % kGridN, iy = cidx / kGridN;
    const double kBig = 1.0e12;
    const bool border = ix == 0 || iy == 0 || ix == kGridN - 1 || iy == kGridN - 1;
    const double x0 = ix == 0 ? -kBig : gx0 + ix * cw, x1 = ix == kGridN - 1 ? kBig : gx0 + (ix + 1) * cw;
    const double y0 = iy == 0 ? -kBig : gy0 + iy * ch, y1 = iy == kGridN - 1 ? kBig : gy0 + (iy + 1) * ch;
    // a finite point of the cell (used when no edge comes near: the cell is then entirely inside or outside)
    const double px = ix == 0 ? x1 : x0, py = iy == 0 ? y1 : y0;
    unsigned masks[2] = {0u, 0u}, flags = 0u;
    for (int b = 0; b < 2; ++b) {
        const double *v = hull_xy + ((size_t)s * 2 + b) * maxv_in * 2;
        const int n = hull_n[s * 2 + b];
        bool touch = false, pin = true;
        for (int i = 0; i < n; ++i) {
            const int j = i == 0 ? n - 1 : i - 1;
            const double ax = v[2 * j], ay = v[2 * j + 1], bx = v[2 * i], by = v[2 * i + 1];
            const double d = seg_rect_dist(ax, ay, bx, by, x0, y0, x1, y1);
            if (d <= reach) masks[b] |= 1u << i;
            if (d <= touch_margin) touch = true;
            // outward normal of a CCW loop is (ey, -ex): p is inside iff it is behind every plane
            if ((by - ay) * (px - bx) - (bx - ax) * (py - by) > 0.0) pin = false;
        }
        if (touch || pin) {
            flags |= 1u << b;
            if (border) masks[b] = n >= 32 ? kFull : ((1u << n) - 1u);      // unbounded cell: the inside test needs every plane
        }
    }
    grid[t] = make_uint4(masks[0], masks[1], flags, 0u);
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

cudaError_t launch_build_grid(const double *hull_xy, const int *hull_n, int n_scen, int maxv_in, double gx0, double gy0,
                              double cw, double ch, double reach, double touch_margin, uint4 *grid, cudaStream_t stream)
{
    const long long total = (long long)n_scen * kGridN * kGridN;
    build_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(hull_xy, hull_n, n_scen, maxv_in, gx0, gy0, cw, ch,
                                                                          reach, touch_margin, grid);
    return cudaGetLastError();
}

cudaError_t launch_stats_reduce(double *slots, double *out, int clear, cudaStream_t stream)
{
    stats_reduce_kernel<<<1, 32 * kStatLen, 0, stream>>>(slots, out, clear);
    return cudaGetLastError();
}

}  // namespace shipsim
