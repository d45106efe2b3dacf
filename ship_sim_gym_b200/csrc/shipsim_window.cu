// shipsim_window.cu -- time-parallel ("window") variant of the fused ShipEnv step kernel, for latency-bound batches.
// sm_100a only.  Same transition as step_kernel (shipsim_kernels.cu), same helper functions
// (shipsim_geom.cuh), bit-identical results (tests/test_gpu_window.py).
//
// Why.  With few envs per SM (4,096 envs = 28 per SM) a K-step rollout is a chain of K dependent iterations and the
// machine idles on instruction latency.  But only the cheap part of a step is sequential: the rigid-body recurrence
// (cpBodyUpdatePosition / cpBodyUpdateVelocity: ~10 flops) and the rudder clamp.  Everything expensive -- the lidar
// query, the ship-vs-bank overlap test, the goal tests, the observation frame -- is a pure function of one pose and
// feeds back into the trajectory only through `done` (auto-reset).  So, when the actions of the rollout are known up
// front (they are: shipsim_step takes the whole [K][N] action tensor), T consecutive steps of one env are SPECULATED:
//   1. the sequential part runs as two short scans executed by every lane of the env's group in lockstep (no
//      divergence, no waiting): rudder / angular velocity / angle from pre-decoded actions, then ONE sincos per lane,
//      then velocity / position with the thrust terms broadcast by shuffles; the group's first lane drops each step's
//      result into shared memory and lane t = lane % T picks up the pose of "its" step;
//   2. lane t does the pose-dependent work of step t -- reach-grid lookup, plane phase, goal tests, out-of-bounds --
//      and the cooperative ray / separating-axis passes run over (env, step) pairs instead of envs;
//   3. the sequential leftovers are resolved without loops over steps where possible: goals taken (one ballot per
//      goal), episode return (ordered per-lane prefix, same rounding as the serial kernel), sticky lidar readings
//      (models.py:71: a miss keeps the last hit -- per ray, the latest hitting lane at or before t, by ballot + clz);
//   4. the window is cut after the first `done`: steps beyond it are discarded, the env resets and the next window
//      starts from the reset state (episodes last ~48 steps on the default map, so a window of 16 commits 13.9 steps
//      on average).
// A warp holds E = 32/T envs; envs of a warp advance independently (each has its own step cursor k0).
//
// Shared memory per env: the plane-phase rows live in a ring of T+1 slots -- slot cb is the "carry", the planes of the
// last committed step (or of the reset), i.e. what step k0's lidar looks at; slot cb+1+t belongs to speculated step t;
// committing n steps just advances cb by n.  The observation frames sit in T+1 linear slots: slot 0 is the carry frame
// (rewritten after every window by the lane of the last committed step), slot 1+t the frame of step t, so that the
// copy-out addresses are compile-time offsets.
//   1. uses a log2(T)-round prefix for the integer rudder recurrence (clamped additions compose), not a T-step chain;
//   2. runs the separating-axis test per env group on the earliest undecided step and, with auto-reset, drops the steps
//      behind the first done; lidar queries are cast for committed steps only.
#include "shipsim_geom.cuh"
#include "shipsim_launch.h"

namespace shipsim {

// The two physics scans are unrolled over the whole window when it is short (every shuffle that feeds the chains is then
// issued ahead of them: 4 -> 16 iterations per trip took the headline from 0.458 to 0.449 ms), by 4 for the longest
// window (T = 32 fully unrolled is 9 % SLOWER at 2,048 envs: code size).
#ifndef SHIPSIM_SCAN_UNROLL
#define SHIPSIM_SCAN_UNROLL (T <= 16 ? T : 4)
#endif
#define SHIPSIM_PRAGMA_(x) _Pragma(#x)
#define SHIPSIM_UNROLL(n) SHIPSIM_PRAGMA_(unroll n)
#ifndef SHIPSIM_WIN_THREADS
#define SHIPSIM_WIN_THREADS 32
#endif
constexpr int kWinThreads = SHIPSIM_WIN_THREADS;      // one-warp CTAs: a latency-bound batch is a few thousand warps, spread them evenly over the SMs (64 -> 32 threads: 0.460 -> 0.458 ms on the headline, -1.8 % at 2,048 envs)
#ifdef SHIPSIM_WIN_MAXNREG
#define SHIPSIM_WIN_BOUNDS __maxnreg__(SHIPSIM_WIN_MAXNREG)
#else
#define SHIPSIM_WIN_BOUNDS __launch_bounds__(kWinThreads, 512 / SHIPSIM_WIN_THREADS)
#endif

template <int T, int HIST>
__global__ void SHIPSIM_WIN_BOUNDS window_kernel(const __grid_constant__ StepParams p)
{
    constexpr int E = 32 / T;                   // envs per warp
    constexpr int NW = kWinThreads / 32;
    constexpr int OBS4 = 4 * HIST;              // float4 per obs row
    constexpr int NS = T + 1;                   // ring slots per env
    constexpr int FS4 = 5;                      // float4 per frame slot (4 used; odd stride: conflict-free)
    constexpr int FR4 = NS * FS4;               // float4 per env frame ring
    constexpr int SC4 = NS * kScr4;             // float4 per env plane-row ring
    constexpr float kMiss = -2.f;               // "this ray did not hit at this step" until the sticky scan resolves it
    __shared__ float4 s_frame[NW * E * FR4];
    __shared__ float4 s_scr[NW * E * SC4];
    __shared__ float2 s_goal[NW * E * kGoals];
    __shared__ float4 s_rf[NW * E * 2];         // reset frame, first two float4 (the rest is -1)
    __shared__ float4 s_ph[NW * 32 * 2];        // physics scan -> lanes: (th, w, bits(rudder), -) and (x, y, vx, vy) per step
    __shared__ int4 s_env[NW * E];              // per env, for the copy-out: step cursor, committed steps, (carry slot of the plane ring), reset?
    __shared__ float s_ray[2 * 32];
    __shared__ unsigned s_src[NW][32];          // ray pass: compacted (plane row | frame offset) of the needy steps
    __shared__ float s_stat[NW][8];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int el = lane / T, t = lane % T;
    const int gbase = el * T;                   // first lane of this env's group
    const unsigned segmask = T == 32 ? kFull : (((1u << T) - 1u) << gbase);
    const int warp_env0 = (blockIdx.x * NW + warp) * E;
    const int e = warp_env0 + el;
    const bool valid = e < p.N;
    const int ec = valid ? e : p.N - 1;         // lanes of idle groups shadow the last env; they never store
    const int wfr0 = warp * E * FR4, wsc0 = warp * E * SC4;
    const int fr0 = wfr0 + el * FR4, sc0 = wsc0 + el * SC4;
    const int goal0 = (warp * E + el) * kGoals;
    const int rf0 = (warp * E + el) * 2;
    const float L = p.lidar_len;
    const long long gid = p.env_id_offset + e;
#define stat (s_stat[warp])
    const unsigned stat_s = smem_addr(s_stat[warp]);
    auto ring = [](int s) { return s >= NS ? s - NS : s; };

    // lane roles in the cooperative passes (as in step_kernel)
    const int rslot = lane / kBeams;
    const int rj = lane - rslot * kBeams;
    if (threadIdx.x < 32) {
        s_ray[threadIdx.x] = p.ray_c[threadIdx.x % kBeams];
        s_ray[32 + threadIdx.x] = p.ray_s[threadIdx.x % kBeams];
    }
    if (lane < 8) stat[lane] = 0.f;
    // separating-axis pass: one step per env group at a time, lane t takes bank edges t, t + T, ...
    const int epl = (p.hull_max + T - 1) / T;

    const size_t act_esize = p.action_dtype == 1 ? 8 : (p.action_dtype == 2 ? 1 : 4);
    const char *act0 = reinterpret_cast<const char *>(p.actions) + (size_t)ec * act_esize;
    const size_t act_stride = (size_t)p.N * act_esize;

    EnvRegs r;
    int cb = 0;                                 // ring slot of the carry
    int k0 = 0;                                 // this env's next step
    bool goals_dirty = false;
    float c0, s0;                               // trig of the pose step k0 starts from
    {
        float4 l0, l1, l2, g0, g1, g2;
        load_env(p, ec, r, l0, l1, l2, g0, g1, g2);
        float2 g[kGoals];
        unpack_goals(g0, g1, g2, g);
        float gx, gy;
        closest_goal(g, r.alive, r.x, r.y, gx, gy);
        sincos_fast(r.th, s0, c0);
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < kGoals; ++i) s_goal[goal0 + i] = g[i];
            float4 *f = s_frame + fr0;          // carry frame = frame of the current state
            f[0] = make_float4(r.x, r.y, (float)r.rudder, r.th);
            f[1] = make_float4(gx, gy, l0.x, l0.y);
            f[2] = make_float4(l0.z, l0.w, l1.x, l1.y);
            f[3] = make_float4(l1.z, l1.w, l2.x, l2.y);
            float hx, hy;                       // carry row = planes of the current pose
            hull_half_extents(p, c0, s0, hx, hy);
            const uint4 cell = load_cell(p, r.scen, r.x + hx, r.y + hy);
            float4 *row = s_scr + sc0;
            if ((cell.x | cell.y | (cell.z & 3u)) != 0u) plane_phase<false, false>(p, r.x, r.y, hx, hy, c0, s0, r.scen, cell, RowPtr{row}, RowPtr{nullptr});
            else row[0] = make_float4(c0, s0, 0.f, 0.f);
        }
    }
    // Actions are held one window ahead: a_my = action of step k0 + t, a_nx = action of step k0 + T + t.  After a window
    // that commits n steps the next window's actions are n lanes further along this pair (two shuffles); only the n
    // lanes at the far end load anything, and what they load is not looked at before the end of the next window.
    // The scenario of the NEXT episode depends only on (seed, env id, episode number): it is drawn an episode ahead, so
    // that at a reset the Philox rounds are not in front of the loads of the new scenario's goals and spawn row, and
    // every window pulls those few lines towards L1 in case it ends in a reset.
    constexpr int SPL = (kScr4 + T - 1) / T;    // spawn-row float4s per lane
    const unsigned pkey = pick_key(p, gid);
    int next_scen = pick_scenario_keyed(p, pkey, r.episode + 1);
    int a_my = 3, a_nx = 3;
    if (valid && t < p.K) a_my = load_action(p, act0 + (size_t)t * act_stride, t, gid);
    if (valid && T + t < p.K) a_nx = load_action(p, act0 + (size_t)(T + t) * act_stride, T + t, gid);
    __syncthreads();

#pragma unroll 1
    while (true) {
        const int nvalid = valid ? min(T, p.K - k0) : 0;        // steps this env still has, capped by the window
        if (__ballot_sync(kFull, nvalid > 0) == 0u) break;
        const bool active = t < nvalid;
        // (Pays only for the longest window, where an SM holds few envs and a reset's loads are fully exposed: at 2,048 envs
        // +6 %; with T <= 16 the twelve prefetches and their addresses cost more than they hide: -1.2 % on the headline.)
        if (T >= 32 && p.auto_reset) {
#pragma unroll
            for (int i = 0; i < (3 + kScr4 + T - 1) / T; ++i) {
                const int q = t + i * T;
                if (q < 3) prefetch_l1(p.bank + (size_t)next_scen * p.scen_stride4 + 2 + q);
                else if (q < 3 + kScr4) prefetch_l1(p.spawn_rows + (size_t)next_scen * kScr4 + (q - 3));
            }
        }

        // ---- 1. the sequential part, T steps: handle_discrete_action (game.py:140-153), cpBodyUpdatePosition,
        // cpBodyUpdateVelocity -- the same operations in the same order as step_kernel, split so that as little as
        // possible is serial.  Every lane of the group runs the scans in lockstep (no divergence, no waiting); the
        // group's first lane drops each step's result into shared memory and lane t picks up step t.
        float mx, my, mth, mc, ms;
        int mrud;
        {
            float4 *ph = s_ph + (warp * 32 + gbase) * 2;
            // scan 1a: the rudder (Ship.rotate / clamp_rudder, models.py:136-146) is an integer recurrence
            // rud <- clamp(rud + inc, -10, 10).  Clamped additions are closed under composition --
            // min(max(r + a, lo), hi) followed by (a', lo', hi') is (a + a', max(lo + a', lo'), min(max(hi + a', lo'), hi'))
            // -- so the T rudder values of the window come from a log2(T)-round prefix over (a, lo, hi) instead of a
            // T-step chain; integers, hence exact.
            int rud_t;                                  // rudder after the action of step t
            {
                int fa = a_my == 1 ? -5 : (a_my == 2 ? 5 : 0), flo = -10, fhi = 10;
#pragma unroll
                for (int off = 1; off < T; off <<= 1) {
                    const int pa = __shfl_up_sync(kFull, fa, off, T), plo = __shfl_up_sync(kFull, flo, off, T),
                              phi = __shfl_up_sync(kFull, fhi, off, T);
                    if (t >= off) {                     // earlier steps first (pa, plo, phi), then this lane's
                        const int na = pa + fa, nlo = max(plo + fa, flo), nhi = min(max(phi + fa, flo), fhi);
                        fa = na; flo = nlo; fhi = nhi;
                    }
                }
                rud_t = min(max(r.rudder + fa, flo), fhi);
            }
            // scan 1b: angular velocity and angle.  Thrust (action 0) leaves the rudder alone, so the torque term of
            // step t uses rud_t; it is rounded to fp32 before the fused multiply-add, exactly as in step_kernel.
            {
                const float dw_t = a_my == 0 ? -p.ang_dt * (float)rud_t : 0.f;
                float th = r.th, w = r.w;
                float2 *po = reinterpret_cast<float2 *>(ph);
SHIPSIM_UNROLL(SHIPSIM_SCAN_UNROLL)
                for (int i = 0; i < T; ++i) {
                    const float dw = __shfl_sync(kFull, dw_t, i, T);
                    th += w * p.dt;
                    w = w * p.damping + dw;
                    if (t == 0) *po = make_float2(th, w);
                    po += 4;
                }
            }
            __syncwarp();
            {
                const float2 v = *reinterpret_cast<const float2 *>(ph + 2 * t);
                mth = v.x; mrud = rud_t;
            }
            sincos_fast(mth, ms, mc);                   // one sincos per lane instead of T per lane
            // thrust of step t acts along the heading the step STARTS from: the trig of step t-1 (or of the carry)
            float pc = __shfl_up_sync(kFull, mc, 1, T), ps = __shfl_up_sync(kFull, ms, 1, T);
            if (t == 0) { pc = c0; ps = s0; }
            float dvx_t = 0.f, dvy_t = 0.f;
            if (a_my == 0) { dvx_t = -p.acc_dt * ps; dvy_t = p.acc_dt * pc; }
            // scan 2: velocity and position
            {
                float x = r.x, y = r.y, vx = r.vx, vy = r.vy;
                float4 *po = ph + 1;
SHIPSIM_UNROLL(SHIPSIM_SCAN_UNROLL)
                for (int i = 0; i < T; ++i) {
                    const float dvx = __shfl_sync(kFull, dvx_t, i, T), dvy = __shfl_sync(kFull, dvy_t, i, T);
                    x += vx * p.dt;
                    y += vy * p.dt;
                    vx = vx * p.damping + dvx;
                    vy = vy * p.damping + dvy;
                    if (t == 0) *po = make_float4(x, y, vx, vy);
                    po += 2;
                }
            }
            __syncwarp();
            {
                const float4 v = ph[2 * t + 1];
                mx = v.x; my = v.y;
            }
        }

        // ---- 2. pose-dependent work of step t at (mx, my, mth): reach-grid cell, candidate planes staged by cp.async
        const int myslot_ring = ring(cb + 1 + t);
        float4 *myrow = s_scr + sc0 + myslot_ring * kScr4;
        float4 *myfr = s_frame + fr0 + (1 + t) * FS4;    // frames: slot 0 = carry, slot 1 + t = step t (no ring)
        float hx = 0.f, hy = 0.f;
        uint4 cell = make_uint4(0u, 0u, 0u, 0xffffffffu);
        if (active) {
            hull_half_extents(p, mc, ms, hx, hy);
            cell = load_cell(p, r.scen, mx + hx, my + hy);
        }

        // goals (collide_goal, game.py:243-257): which goals does the hull touch at this pose?  (whether they are
        // still alive is settled by the scan below)
        float2 g[kGoals];
        float gd2[kGoals];
        unsigned touch = 0u, cand = 0u;
        {
#pragma unroll
            for (int i = 0; i < kGoals; ++i) {
                g[i] = s_goal[goal0 + i];
                const float ux = g[i].x - mx, uy = g[i].y - my;
                gd2[i] = ux * ux + uy * uy;
            }
#pragma unroll
            for (int i = 0; i < kGoals; ++i)
                if (active && ((r.alive >> i) & 1) && gd2[i] <= p.goal_cull_r2) cand |= 1u << i;
        }
        // this step's lidar slots: "no hit" until the ray pass says otherwise
        reinterpret_cast<float2 *>(myfr + 1)[1] = make_float2(kMiss, kMiss);
        myfr[2] = make_float4(kMiss, kMiss, kMiss, kMiss);
        myfr[3] = make_float4(kMiss, kMiss, kMiss, kMiss);

        const bool near_any = active && (cell.x | cell.y | (cell.z & 3u)) != 0u;
        const bool staged = near_any && ((cell.z >> 8) & 0xffu) <= (unsigned)kMaxCand;
        if (staged) {
            const float4 *E4 = reinterpret_cast<const float4 *>(p.edges_d + (size_t)r.scen * (2 * kMaxHull));
#pragma unroll
            for (int n = 0; n < kMaxCand; ++n) {
                const unsigned idx = (cell.w >> (8 * n)) & 0xffu;
                if (idx != 0xffu) {
                    cp_async16(myrow + 1 + 2 * n, E4 + 2 * idx);
                    cp_async16(myrow + 2 + 2 * n, E4 + 2 * idx + 1);
                }
            }
        }
        const bool oob = (mx < 0.f) || (mx > p.W) || (my < 0.f) || (my > p.H);
        // (while the candidate planes are in flight: the exact goal tests and everything that follows from them)
#pragma unroll 1
        while (cand) {
            const int i = __ffs(cand) - 1;
            cand &= cand - 1u;
            const float2 gi = s_goal[goal0 + i];
            const float ux = gi.x - mx, uy = gi.y - my;
            const float qx = fmaf(ux, mc, __fmul_rn(uy, ms)), qy = fmaf(uy, mc, -__fmul_rn(ux, ms));
            if (goal_contact(p, qx, qy)) touch |= 1u << i;
        }

        // ---- 3a. goals taken so far in the window: prefix OR of the touch masks (one ballot per goal: taken before /
        // up to step t <=> an earlier / this-or-earlier lane of the env touches it)
        unsigned inc = 0u, exc = 0u;
        {
            const unsigned below = segmask & ((1u << lane) - 1u), upto = segmask & ((2u << lane) - 1u);
#pragma unroll
            for (int i = 0; i < kGoals; ++i) {
                const unsigned m = __ballot_sync(kFull, (touch >> i) & 1u);
                if (m & below) exc |= 1u << i;
                if (m & upto) inc |= 1u << i;
            }
        }
        const int alive_prev = r.alive & ~(int)exc, alive_t = r.alive & ~(int)inc;
        const bool goal_reached = (alive_prev & (int)touch) != 0;
        const int steps_t = r.steps + t + 1;
        const bool all_goals = alive_t == 0;
        const bool timeout = steps_t >= p.max_steps;
        // With auto-reset the window is cut after the env's first done step, so nothing behind the first step that is done
        // for a reason already known (no goals left, out of bounds, step cap) can be committed: such steps need neither
        // the overlap test nor a lidar query.  (A ship that has run into a bank keeps ploughing through it for the rest of
        // the speculated window: before this cut those steps were a third of the separating-axis work.)
        const bool prune = p.auto_reset != 0;
        bool open_step = true;
        {
            const unsigned cd = __ballot_sync(kFull, active && (all_goals || oob || timeout)) & segmask;
            if (prune && cd != 0u && lane > __ffs(cd) - 1) open_step = false;
        }


        // ---- plane phase at pose t: lidar planes of step t+1 + ship-vs-bank pre-test of step t
        cp_async_wait_all();
        unsigned ask = 0u;
        if (staged) {
            ask = plane_phase<true, true>(p, mx, my, hx, hy, mc, ms, r.scen, cell, RowPtr{myrow}, RowPtr{myrow + 1});
        } else if (near_any) {                  // more candidates than are staged: evaluated straight from global memory
            ask = plane_phase_unstaged(p, mx, my, hx, hy, mc, ms, r.scen, cell, myrow);
        } else if (active) {
            myrow[0] = make_float4(mc, ms, 0.f, 0.f);
        }
        __syncwarp();

        // ---- overlap test at the new pose -> collide_ship (game.py:232-241): separating-axis pass for the steps the
        // plane phase could not settle.  Every env group works on its own earliest open step (one lane per bank edge,
        // strided when the hull has more edges than the group has lanes); a collision closes all later steps of the env.
        bool colliding = false;
        bool asking = ask != 0u && open_step;
        {
            unsigned needs;
            while ((needs = __ballot_sync(kFull, asking)) != 0u) {
                const unsigned mine = needs & segmask;
                const bool act_env = mine != 0u;
                const int src = act_env ? __ffs(mine) - 1 : lane;
                const float bx = __shfl_sync(kFull, mx, src), by = __shfl_sync(kFull, my, src);
                const float bc = __shfl_sync(kFull, mc, src), bs = __shfl_sync(kFull, ms, src);
                const unsigned bask = __shfl_sync(kFull, ask, src);
                const float4 *rec = p.bank + (size_t)r.scen * p.scen_stride4;
                const float4 hdr = __ldg(rec + 4);
                const float4 *bE = rec + kBankHeader4;
                float rx[kShipVerts], ry[kShipVerts];
#pragma unroll
                for (int j = 0; j < kShipVerts; ++j) {
                    rx[j] = p.ship_lx[j] * bc - p.ship_ly[j] * bs;
                    ry[j] = p.ship_lx[j] * bs + p.ship_ly[j] * bc;
                }
                bool coll = false;
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    const bool do_b = act_env && ((bask >> b) & 1u) && !coll;
                    const int nb = __float_as_int(b ? hdr.w : hdr.z);
                    const float4 *bEb = bE + b * p.maxv;
                    // a bank edge normal separates?  (lane t of the group takes edges t, t + T, ...)
                    bool lsep = false;
#pragma unroll 1
                    for (int u = 0; u < epl; ++u) {
                        const int idx = t + u * T;
                        if (do_b && idx < nb) lsep = lsep || bank_axis_separates(__ldg(bEb + idx), rx, ry, bx, by);
                    }
                    const unsigned sb = __ballot_sync(kFull, lsep);
                    bool sep = (sb & segmask) != 0u;
                    if (__ballot_sync(kFull, do_b && !sep)) {                   // else try the ship's edge normals
                        float pr[kShipVerts];
#pragma unroll
                        for (int j = 0; j < kShipVerts; ++j) pr[j] = 3.0e38f;
#pragma unroll 1
                        for (int u = 0; u < epl; ++u) {
                            const int idx = t + u * T;
                            if (do_b && idx < nb) {
                                const float4 ed = __ldg(bEb + idx);
#pragma unroll
                                for (int j = 0; j < kShipVerts; ++j) {
                                    const float nx = p.ship_nx[j] * bc - p.ship_ny[j] * bs;
                                    const float ny = p.ship_nx[j] * bs + p.ship_ny[j] * bc;
                                    pr[j] = fminf(pr[j], nx * (ed.z - bx) + ny * (ed.w - by));
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < kShipVerts; ++j) {      // axis j separates <=> no bank vertex reaches the hull's plane j
                            const unsigned reach = __ballot_sync(kFull, pr[j] <= p.ship_off[j]);
                            sep = sep || (reach & segmask) == 0u;
                        }
                    }
                    if (do_b && !sep) coll = true;
                }
                if (act_env && lane == src) { asking = false; colliding = coll; }
                if (coll && prune) asking = false;                              // the env is done at step src: later steps are moot
            }
        }

        // ShipEnv.determine_reward (ship_env.py:62-77) / is_done (ship_env.py:115-134)
        const float reward = goal_reached ? 1.f : (oob ? -1.f : p.step_penalty);
        const bool done = colliding || all_goals || oob || timeout;
        // cut the window after the first done step (auto-reset): everything speculated beyond it is discarded
        const unsigned dmask = __ballot_sync(kFull, active && done) & segmask;
        const bool do_reset = p.auto_reset && dmask != 0u;
        const int ncommit = do_reset ? __ffs(dmask) - gbase : nvalid;
        const bool commit = t < ncommit;

        // The actions of the next window, now: after a window that commits n steps they are n lanes further along the
        // (a_my, a_nx) pair, and the lanes at the far end load theirs.  Done here, as soon as n is known, and not at the end
        // of the iteration: the compiler waits for every outstanding load at the loop's back edge, so a load issued there
        // was paid in full at the top of every window (4.5 % of the stall samples, ncu); from here it has the ray pass, the
        // frame and the copy-out to land behind.
        {
            const int sh = t + ncommit;                             // (ncommit <= T)
            const int m1 = __shfl_sync(kFull, a_my, sh, T), m2 = __shfl_sync(kFull, a_nx, sh, T);    // lane (sh mod T) of the group
            a_my = sh < T ? m1 : m2;
            a_nx = m2;
            if (sh >= T) {
                const int kk = k0 + ncommit + T + t;
                a_nx = (valid && kk < p.K) ? load_action(p, act0 + (size_t)kk * act_stride, kk, gid) : 3;
            }
            // two windows ahead: pull the action rows into L2 (longest window only: -1.2 % on the headline with T = 16)
            if (T >= 32 && p.action_dtype != 3 && valid && k0 + ncommit + 2 * T + t < p.K) prefetch_l2(act0 + (size_t)(k0 + ncommit + 2 * T + t) * act_stride);
        }


        // ---- LiDAR.query (models.py:39-76) of step t at its PRE-integration pose = the pose of step t-1 (ring slot
        // cb + t; the carry for t = 0), for the steps that are committed.  Up to three needy steps per pass, lanes 0-9 /
        // 10-19 / 20-29 = their rays.
        {
            const int srcrow = sc0 + ring(cb + t) * kScr4;
            const int hz_own = commit ? __float_as_int(s_scr[srcrow].z) : 0;
            const bool big = (hz_own & kHdrBig) != 0;
            const bool wants = (hz_own & 0x3ff) != 0;
            const unsigned need = __ballot_sync(kFull, wants);
            if (big) ray_query_serial(p, s_scr + srcrow, reinterpret_cast<float *>(myfr) + 6);
            if (need) {
                if (wants) s_src[warp][__popc(need & ((1u << lane) - 1u))] =
                    (unsigned)(srcrow - wsc0) | ((unsigned)((fr0 - wfr0 + (1 + t) * FS4) * 4 + 6) << 16);
                __syncwarp();
                const int cnt = __popc(need);
                const float ray_c = s_ray[lane], ray_s = s_ray[32 + lane];
#pragma unroll 1
                for (int q = rslot; q < cnt; q += 3) {
                    if (rslot < 3) {
                        const unsigned sw = s_src[warp][q];
                        const int rw = wsc0 + (int)(sw & 0xffffu);
                        const float4 hdr = s_scr[rw];
                        float dirx, diry;
                        ray_dir(hdr.x, hdr.y, ray_c, ray_s, dirx, diry);
                        const int hz = __float_as_int(hdr.z);
                        const int n = hz & 0xff;
                        float v0 = -1.f, v1 = -1.f;
#pragma unroll 1
                        for (int i = 0; i < n; ++i) {
                            const float4 e0 = s_scr[rw + 1 + 2 * i];
                            const float2 e1 = *reinterpret_cast<const float2 *>(s_scr + rw + 2 + 2 * i);
                            float val;
                            const bool ok = ray_vs_plane(e0.x, e0.y, e0.z, e0.w, e1.x, dirx, diry, L, val);
                            if (ok && e1.y == 0.f) v0 = val;
                            if (ok && e1.y != 0.f) v1 = val;
                        }
                        if (hz & kHdrIn0) v0 = L;
                        if (hz & kHdrIn1) v1 = L;
                        const float v = v0 >= 0.f ? v0 : v1;
                        if (v >= 0.f) reinterpret_cast<float *>(s_frame + wfr0)[(sw >> 16) + rj] = v;
                    }
                }
            }
        }

        // episode return after step t: ordered sum (lane t adds the rewards of steps 0..t one by one), same rounding
        // as one step at a time
        float my_ret = r.ret;
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const float rv = __shfl_sync(kFull, reward, i, T);
            if (i <= t) my_ret += rv;
        }
        warp_stats(stat_s, lane, commit, goal_reached, done, colliding, oob, timeout, all_goals, my_ret, steps_t);
        // this step's frame: pose, rudder, nearest remaining goal (closest_goal, game.py:333-349), and the lidar
        // readings.  Sticky readings (models.py:71): a ray that missed keeps the reading of the last step at which it
        // hit -- the latest hit at or before step t inside the window (found with a ballot per ray), else the carry's.
        __syncwarp();                           // ray pass results are in the frame slots
        {
            float gx = -1.f, gy = -1.f, best = 3.0e38f;
#pragma unroll
            for (int i = 0; i < kGoals; ++i)
                if (((alive_t >> i) & 1) && gd2[i] < best) { best = gd2[i]; gx = g[i].x; gy = g[i].y; }
            float h[kBeams], cv[kBeams];
            {
                const float4 *cf = s_frame + fr0;
                const float2 a0 = reinterpret_cast<const float2 *>(cf + 1)[1];
                const float4 a1 = cf[2], a2 = cf[3];
                cv[0] = a0.x; cv[1] = a0.y; cv[2] = a1.x; cv[3] = a1.y; cv[4] = a1.z; cv[5] = a1.w;
                cv[6] = a2.x; cv[7] = a2.y; cv[8] = a2.z; cv[9] = a2.w;
                const float2 b0 = reinterpret_cast<const float2 *>(myfr + 1)[1];
                const float4 b1 = myfr[2], b2 = myfr[3];
                h[0] = b0.x; h[1] = b0.y; h[2] = b1.x; h[3] = b1.y; h[4] = b1.z; h[5] = b1.w;
                h[6] = b2.x; h[7] = b2.y; h[8] = b2.z; h[9] = b2.w;
            }
            const unsigned upto = segmask & ((2u << lane) - 1u);       // lanes of this env at steps <= t
#pragma unroll
            for (int j = 0; j < kBeams; ++j) {
                const unsigned m = __ballot_sync(kFull, h[j] != kMiss) & upto;
                const float v = __shfl_sync(kFull, h[j], 31 - __clz(m));     // m == 0: lane 31's value, not used
                h[j] = m ? v : cv[j];
            }
            myfr[0] = make_float4(mx, my, (float)mrud, mth);
            myfr[1] = make_float4(gx, gy, h[0], h[1]);
            myfr[2] = make_float4(h[2], h[3], h[4], h[5]);
            myfr[3] = make_float4(h[6], h[7], h[8], h[9]);
        }

        // ---- 4. the state the next window starts from: after the last committed step (the scans left it in shared
        // memory), or the reset
        const int srcl = gbase + max(ncommit, 1) - 1;
        EnvRegs rn;
        {
            const float4 *ph = s_ph + (warp * 32 + srcl) * 2;
            const float4 v0 = ph[0], v1 = ph[1];
            rn.th = v0.x; rn.w = v0.y;
            rn.x = v1.x; rn.y = v1.y; rn.vx = v1.z; rn.vy = v1.w;
        }
        rn.rudder = __shfl_sync(kFull, mrud, srcl);
        rn.ret = __shfl_sync(kFull, my_ret, srcl);
        rn.alive = __shfl_sync(kFull, alive_t, srcl);
        rn.steps = r.steps + ncommit; rn.scen = r.scen; rn.episode = r.episode;
        float c0n = __shfl_sync(kFull, mc, srcl), s0n = __shfl_sync(kFull, ms, srcl);
        if (t == 0) s_env[warp * E + el] = make_int4(k0, ncommit, cb, (int)do_reset);
        float4 spv[SPL];                                            // the new scenario's spawn row, lane t holds entries t, t + T, ...
        if (do_reset) {                                             // ShipEnv.reset (ship_env.py:171-184)
            const int ep = r.episode + 1;
            reset_env(p, rn, next_scen, ep);
            c0n = 1.f; s0n = 0.f;
            const float4 *rec = p.bank + (size_t)rn.scen * p.scen_stride4;
            const float4 *sp = p.spawn_rows + (size_t)rn.scen * kScr4;
            const float4 rg0 = __ldg(rec + 2), rg1 = __ldg(rec + 3), rg2 = __ldg(rec + 4);
#pragma unroll
            for (int i = 0; i < SPL; ++i) if (t + i * T < kScr4) spv[i] = __ldg(sp + t + i * T);
            next_scen = pick_scenario_keyed(p, pkey, ep + 1);
            float2 gn[kGoals];
            unpack_goals(rg0, rg1, rg2, gn);
            float rgx, rgy;
            closest_goal(gn, rn.alive, rn.x, rn.y, rgx, rgy);
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < kGoals; ++i) s_goal[goal0 + i] = gn[i];
                s_rf[rf0] = make_float4(rn.x, rn.y, 0.f, 0.f);
                s_rf[rf0 + 1] = make_float4(rgx, rgy, -1.f, -1.f);
            }
            goals_dirty = true;
        }
        __syncwarp();

        // ---- outputs of the committed steps.  Obs row of step t = [frame t-1 | frame t] = frame slots (t, t + 1);
        // OBS4 lanes per row, 32/OBS4 rows per store instruction, 128-bit streaming stores, no loop: a window has at
        // most T rows per env.
        if (p.obs) {
            constexpr int RPR = 32 / OBS4;                          // rows per store instruction
            constexpr int NIT = (T + RPR - 1) / RPR;
            const float4 neg = make_float4(-1.f, -1.f, -1.f, -1.f);
            const int col = lane % OBS4, r0 = lane / OBS4;
            const int half = HIST == 2 ? (col >> 2) : 1, q = col & 3;
            const size_t rstride = (size_t)p.N * OBS4 * RPR;        // float4 between the rows of consecutive stores
#pragma unroll
            for (int elr = 0; elr < E; ++elr) {
                const int4 ev = s_env[warp * E + elr];              // k0, committed, -, reset
                const float4 *fb = s_frame + wfr0 + elr * FR4 + (r0 + half) * FS4 + q;
                float4 *o = p.obs + ((size_t)(ev.x + r0) * p.N + (warp_env0 + elr)) * OBS4 + col;
                const int nplain = ev.w ? ev.y - 1 : ev.y;          // the last committed row of an env that resets is special
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    if (r0 + it * RPR < nplain) __stcs(o, fb[it * RPR * FS4]);
                    o += rstride;
                }
                if (ev.w && r0 == 0) {                              // ship_env.py:180-184: [-1 x 16 | reset frame], vals = -1
                    const float4 rfv = (half == 0 || q >= 2) ? neg : s_rf[(warp * E + elr) * 2 + q];
                    __stcs(p.obs + ((size_t)(ev.x + nplain) * p.N + (warp_env0 + elr)) * OBS4 + col, rfv);
                }
            }
        }
        if (commit) {
            const size_t row = (size_t)(k0 + t) * p.N + e;
            if (p.reward) p.reward[row] = reward;
            if (p.done) p.done[row] = done ? 1 : 0;
        }
        __syncwarp();                           // rows and frames have been read: the carry may move

        if (ncommit > 0) {
            const int ncb = ring(cb + ncommit);
            const float4 negc = make_float4(-1.f, -1.f, -1.f, -1.f);
            if (do_reset) {                     // carry := reset frame + the spawn pose's planes (built at scenario upload)
                if (t == 0) {
                    float4 *f = s_frame + fr0;
                    f[0] = s_rf[rf0]; f[1] = s_rf[rf0 + 1]; f[2] = negc; f[3] = negc;
                }
                float4 *row = s_scr + sc0 + ncb * kScr4;            // the whole row: entries past its planes are never read
#pragma unroll
                for (int i = 0; i < SPL; ++i) if (t + i * T < kScr4) row[t + i * T] = spv[i];
            } else if (t == ncommit - 1) {      // carry := the frame of the last committed step
                float4 *f = s_frame + fr0;
                f[0] = myfr[0]; f[1] = myfr[1]; f[2] = myfr[2]; f[3] = myfr[3];
            }
            cb = ncb;
            r = rn; c0 = c0n; s0 = s0n;
            k0 += ncommit;
        }
        __syncwarp();
    }

    if (valid && t == 0) {
        const float4 *f = s_frame + fr0;
        const float4 l1 = f[1], l2 = f[2], l3 = f[3];
        store_env(p, e, r, make_float4(l1.z, l1.w, l2.x, l2.y), make_float4(l2.z, l2.w, l3.x, l3.y), l3.z, l3.w);
        if (goals_dirty) {
            const float2 a0 = s_goal[goal0], a1 = s_goal[goal0 + 1], a2 = s_goal[goal0 + 2], a3 = s_goal[goal0 + 3], a4 = s_goal[goal0 + 4];
            store_goals(p, e, make_float4(a0.x, a0.y, a1.x, a1.y), make_float4(a2.x, a2.y, a3.x, a3.y), make_float4(a4.x, a4.y, 0.f, 0.f));
        }
    }
    __syncwarp();
    if (lane < 8 && p.stats) {
        const float v = stat[lane];
        if (v != 0.f) atomicAdd(p.stats + (size_t)(blockIdx.x % kStatSlots) * kStatLen + lane, (double)v);
    }
#undef stat
}

template <int T>
static cudaError_t launch_t(const StepParams &p, cudaStream_t stream, LaunchShape *shape)
{
    const int envs_per_cta = (kWinThreads / 32) * (32 / T);
    const int blocks = (p.N + envs_per_cta - 1) / envs_per_cta;
    if (shape) { shape->lanes_per_env = T; shape->threads = kWinThreads; shape->blocks = blocks; shape->window = T; }
    if (p.history == 2) window_kernel<T, 2><<<blocks, kWinThreads, 0, stream>>>(p);
    else window_kernel<T, 1><<<blocks, kWinThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_window(const StepParams &p, int window, cudaStream_t stream, LaunchShape *shape)
{
    switch (window) {
        case 4: return launch_t<4>(p, stream, shape);
        case 8: return launch_t<8>(p, stream, shape);
        case 16: return launch_t<16>(p, stream, shape);
        case 32: return launch_t<32>(p, stream, shape);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace shipsim
