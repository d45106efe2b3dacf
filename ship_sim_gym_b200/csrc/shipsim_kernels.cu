// shipsim_kernels.cu -- the fused ShipEnv step kernel (K env-steps per launch), reset, stats and bank-build kernels.
// sm_100a only.  See shipsim_device.cuh for the data layout and the reference lines each piece restates.
//
// Execution model.  G lanes cooperate on one env (G = 1..32, 32/G envs per warp); the lanes of a group hold
// identical copies of the env's scalar state.  Per-env work that is regular runs per lane; the irregular geometry is
// done in two ways that keep control flow warp-uniform no matter how few envs of a warp are near a bank:
//   * plane phase (owner lane): the reach grid names the few bank edges near the env; their planes are evaluated in
//     double at the new pose, which at once (a) settles the ship-vs-bank test for all but touching cases and (b) leaves
//     in shared memory everything the NEXT step's lidar needs;
//   * ray pass (whole warp): three needy envs at a time, one lane per (env, ray), reading those planes;
//   * SAT pass (whole warp, rare): 32/lps envs at a time, one lane per (env, bank edge).
// The loop is rotated: an iteration starts with the pose already integrated, so that the grid cell of the pose an
// iteration works on was requested a whole iteration earlier.
#include "shipsim_geom.cuh"
#include "shipsim_launch.h"

#ifndef SHIPSIM_MIN_BLOCKS
#define SHIPSIM_MIN_BLOCKS 4
#endif
// Grids that fill the machine (one lane per env, >= 1,184 CTAs of 128 envs' worth): CTA size and CTAs per SM to aim
// for.  Registers per thread follow from the pair: 128 x 5 -> 96, 64 x 9 -> 112, 128 x 4 / 64 x 8 -> 128.  Spills are
// ruinous here (the L1 that would catch them is almost entirely carved out as shared memory), so the shape is chosen
// as the most warps per SM that still compile without any.
#ifndef SHIPSIM_SMALL_THREADS
#define SHIPSIM_SMALL_THREADS 64
#endif
#ifndef SHIPSIM_BIG_THREADS
#define SHIPSIM_BIG_THREADS 128
#endif
#ifndef SHIPSIM_BIG_MIN_BLOCKS
#define SHIPSIM_BIG_MIN_BLOCKS 5
#endif
// SHIPSIM_TMA_STORE=1: observation rows leave the SM by TMA bulk copies (one cp.async.bulk.global.shared::cta per env
// row, fence.proxy.async + bulk wait_group) instead of the loop of 128-bit shared loads + streaming stores.  Correct
// (the whole GPU suite passes with it) but measured 15-17 % SLOWER (1M envs x 32 steps: 1.68 -> 1.98 ms; hard map 65,536
// envs: 0.52 -> 0.61 ms; profiles/r02_tma_store_ab.log): a 128-byte row is too small a unit for the TMA engine -- 640
// bulk operations per SM per step against 8 LDS + 8 STG per lane that issue in the shadow of other warps.  Off.
#ifndef SHIPSIM_TMA_STORE
#define SHIPSIM_TMA_STORE 0
#endif

namespace shipsim {

constexpr float kDeadGoal = 3.0e19f;        // coordinate of a goal that has been taken: its squared distance is +inf

// MINB = CTAs per SM the register allocation aims for: 5 (96 registers) pays off only when the grid is large enough
// to fill them, otherwise 4 (128 registers, less rematerialisation).
//
// Shared memory: ONE array; every env has one block of EB4 float4 (odd stride: the 128-bit accesses of a warp's envs
// are conflict-free) holding its observation tile row, its plane row and its goals at compile-time offsets, addressed
// from a single per-lane byte offset.  (Round 1 kept seven arrays with their own index arithmetic; at 96 registers the
// compiler rebuilt those indices from threadIdx / blockIdx inside the loop: 12 % of the instructions, and an S2R
// stall that was the hottest sample of the kernel.  The offsets are laundered through an empty asm so that they stay
// in their registers.)
template <int G, int HIST, int MINB, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int EPW = 32 / G;                 // envs per warp
    constexpr int OBS4 = 4 * HIST;              // float4 per obs row
    // ray pass: two passes (six envs) per loop trip in the 128-register build (grids that do not fill the machine are
    // latency bound: +4 % on the hard map at 65,536 envs); the 96-register build has no room for it (-1 % at 1M envs)
    constexpr int kRayPassUnroll = MINB * THREADS <= 512 ? 2 : 1;       // (512 threads per SM = the 128-register builds)
    constexpr int CF = 16 * (HIST - 1);         // float offset of the newest frame inside the tile row
    constexpr int NW = THREADS / 32;
    // env block: [0, OBS4) tile row = [older frame | newest frame] (lidar slots = the sticky readings);
    // [SCR, SCR + kScr4) plane row; [GOL, GOL + 3) goals (taken goals hold kDeadGoal)
    constexpr int SCR = OBS4, GOL = OBS4 + kScr4;
    // [RAW, RAW + 2 * kMaxCand): raw candidate-plane records (cp.async targets) -- only with G > 1 (few envs per CTA),
    // where they are requested at the top of the iteration; with G = 1 that would cost 16 KB, so they land in the plane
    // row itself once the ray pass has finished with it
    constexpr int RAW = GOL + 3;
    constexpr int EB4 = (RAW + (G > 1 ? 2 * kMaxCand : 0)) | 1;
    // warp block: the env blocks, then 2 float4 of ray-pass list, 2 of statistics, 5 of ray directions (cos[10], sin[10])
    constexpr int WX4 = EPW * EB4;
    constexpr int SRC = WX4, STA = WX4 + 2, RAY = WX4 + 4;
    constexpr int WB4 = WX4 + 9;
    // the last float4 of an env block exists only to make the stride odd: it is the landing slot of fetch_cell_async
    constexpr int CEL = EB4 - 1;
    static_assert(CEL >= RAW + (G > 1 ? 2 * kMaxCand : 0), "the cell slot must not overlap the block's contents");
    __shared__ float4 smem[NW * WB4];

    // Everything a lane needs to find its data -- lane, warp, shared-memory offsets, output indices -- follows from ONE
    // value, the env index e, by a mask, a shift or a multiply-add; e itself is laundered (a shuffle from the own lane:
    // an identity the assembler cannot see through) so that nothing is rebuilt from the special registers inside the loop.
    int e = (blockIdx.x * NW + (threadIdx.x >> 5)) * EPW + (threadIdx.x & 31) / G;
    int lane;
    if (G == 1) {
        e = __shfl_sync(kFull, e, threadIdx.x & 31);
        lane = e & 31;
    } else {
        lane = threadIdx.x & 31;
        lane = __shfl_sync(kFull, lane, lane);
        e = __shfl_sync(kFull, e, lane);
    }
    const int grp = lane / G, gl = lane % G;
    // Shared memory is addressed by 32-bit shared-space addresses held in ordinary registers (see "shared memory by
    // address" in shipsim_device.cuh): eb = this env's block (laundered like e), wb = this warp's block, one
    // multiply-add away.
    unsigned eb = smem_addr(smem) + (unsigned)(((e / EPW) & (NW - 1)) * WB4 + grp * EB4) * 16u;   // (blockIdx.x * NW is a multiple of NW)
    eb = __shfl_sync(kFull, eb, lane);
    const unsigned wb = eb - (unsigned)(grp * EB4) * 16u;
    const bool valid = e < p.N;
    const bool leader = valid && gl == 0;
#define EA(i) (eb + 16u * (unsigned)(i))                     /* address of float4 i of this env's block */
#define WA(env, i) (wb + (unsigned)(env) * (EB4 * 16u) + 16u * (unsigned)(i))
    const float L = p.lidar_len;

    if (lane < kBeams) {                                            // ray direction table (body frame), one per warp
        sts1(wb + RAY * 16 + 4 * lane, p.ray_c[lane]);
        sts1(wb + RAY * 16 + 40 + 4 * lane, p.ray_s[lane]);
    }
    if (lane < 8) sts1(wb + STA * 16 + 4 * lane, 0.f);
    if (!valid && gl == 0) sts4(EA(SCR), make_float4(1.f, 0.f, 0.f, 0.f));      // idle groups: a defined heading for the shadow lanes

    EnvRegs r;
    {
        float4 l0, l1, l2, g0, g1, g2;
        load_env(p, valid ? e : p.N - 1, r, l0, l1, l2, g0, g1, g2);   // lanes of idle groups shadow the last env; they never store
        float2 g[kGoals];
        unpack_goals(g0, g1, g2, g);
        float gx, gy;
        closest_goal(g, r.alive, r.x, r.y, gx, gy);
        if (gl == 0) {
#pragma unroll
            for (int i = 0; i < kGoals; ++i) if (!((r.alive >> i) & 1)) g[i] = make_float2(kDeadGoal, kDeadGoal);
            sts4(EA(GOL), make_float4(g[0].x, g[0].y, g[1].x, g[1].y));
            sts4(EA(GOL + 1), make_float4(g[2].x, g[2].y, g[3].x, g[3].y));
            // .z, .w: the goals-alive mask and the episode number live here, not in registers (a taken goal and a reset are
            // the only things that touch them)
            sts4(EA(GOL + 2), make_float4(g[4].x, g[4].y, __int_as_float(r.alive), __int_as_float(r.episode)));
            // newest frame of the resident tile = frame of the current state
            sts4(EA(OBS4 - 4), make_float4(r.x, r.y, (float)r.rudder, r.th));
            sts4(EA(OBS4 - 3), make_float4(gx, gy, l0.x, l0.y));
            sts4(EA(OBS4 - 2), make_float4(l0.z, l0.w, l1.x, l1.y));
            sts4(EA(OBS4 - 1), make_float4(l1.z, l1.w, l2.x, l2.y));
        }
    }
    float c, s, hx, hy;
    sincos_fast(r.th, s, c);
    hull_half_extents(p, c, s, hx, hy);
    // (G > 1: the lanes of a group share the env's block.  One lane fetches / writes, and wherever another lane's read and
    // the leader's write of the same word are not already separated by a warp synchronisation, one is placed -- all of them
    // compile away for G = 1.  compute-sanitizer racecheck: profiles/r02_x_sanitize.txt)
    if (G == 1 || gl == 0) fetch_cell_async(p, r.scen, r.x + hx, r.y + hy, EA(CEL));
    __syncwarp();                                // ray table, statistics and tile initialisation (all per warp) are in place

    // Iteration k >= 0 is env-step k and starts with the pose ALREADY integrated (cpBodyUpdatePosition of step k);
    // iteration -1 only runs the plane phase at the loaded pose and the first integration.
#pragma unroll 1
    for (int k = -1; k < p.K; ++k) {
        const bool live = k >= 0;
        bool done = false, do_reset = false, goal_reached = false;
        float reward = 0.f, gx = -1.f, gy = -1.f;
        if (live) {
            // this step's action: asked for first, looked at after the lidar pass (kept in a register across the loop it was
            // the one value the 96-register build spilled -- and a spilled prefetch is a synchronous load)
            const int a = load_action_at(p, (size_t)k * p.N + min(e, p.N - 1), k, p.env_id_offset + e);   // (idle lanes read the last env's)
            // previous frame <- newest frame of the last step / reset (SURVEY.md App. A note N2); lidar stays in place
            if (SHIPSIM_TMA_STORE && k > 0) bulk_store_wait_read();      // last step's row has left the tile: it may be rewritten
            if (HIST == 2 && gl == 0) {
                const float4 f0 = lds4(EA(4)), f1 = lds4(EA(5)), f2 = lds4(EA(6)), f3 = lds4(EA(7));
                sts4(EA(0), f0); sts4(EA(1), f1); sts4(EA(2), f2); sts4(EA(3), f3);
            }

            // ---- LiDAR.query (models.py:39-76) at the PRE-integration pose (game.py:193 precedes :194): the plane
            // rows hold that pose's planes.  Up to three needy envs per pass, lanes 0-9 / 10-19 / 20-29 = their rays.
            const float4 h0 = lds4(EA(SCR));    // header of the pose the step starts from: its trig, and what its lidar needs
            const int hz_own = __float_as_int(h0.z);
            const bool big = leader && (hz_own & kHdrBig);
            const bool wants = leader && (hz_own & 0x3ff) != 0;
            const unsigned need = __ballot_sync(kFull, wants);
            if (HIST == 2) __syncwarp();        // the frame copy has read the old readings before any lane overwrites them
            if (big) {                          // (rare: generic pointers are good enough)
                float4 *blk = smem + ((e / EPW) & (NW - 1)) * WB4 + grp * EB4;
                ray_query_serial(p, blk + SCR, reinterpret_cast<float *>(blk) + CF + 6);
            }
            if (need) {
                // needy envs, compacted: entry q of the warp's list = the env slot of the q-th needy env
                if (wants) sts_u8(wb + SRC * 16 + __popc(need & ((1u << lane) - 1u)), (unsigned)grp);
                __syncwarp();
                const int cnt = __popc(need);
                const int rslot = lane / kBeams, rj = lane - rslot * kBeams;    // env slot 0..2 of the pass (lanes 30, 31 idle), ray
                const float ray_c = lds1(wb + RAY * 16 + 4 * rj), ray_s = lds1(wb + RAY * 16 + 40 + 4 * rj);     // (lanes of a ray: broadcast)
#pragma unroll kRayPassUnroll
                for (int q = rslot; q < cnt; q += 3) {          // lanes 30, 31 (rslot 3) only keep the others company
                    if (rslot < 3) {
                        const unsigned ra = WA(lds_u8(wb + SRC * 16 + q), 0);      // the env's block
                        const float4 hdr = lds4(ra + SCR * 16);
                        float dirx, diry;
                        ray_dir(hdr.x, hdr.y, ray_c, ray_s, dirx, diry);
                        const int hz = __float_as_int(hdr.z);
                        const int n = hz & 0xff;
                        float v0 = -1.f, v1 = -1.f;                     // hit distance per bank (< 0: none)
#pragma unroll 1
                        for (int i = 0; i < n; ++i) {   // cpPolyShapeSegmentQuery: later accepted edges overwrite earlier ones
                            const float4 e0 = lds4(ra + (SCR + kRowPlane0 + 2 * i) * 16);
                            const float2 e1 = lds2(ra + (SCR + kRowPlane0 + 1 + 2 * i) * 16);
                            float val;
                            const bool ok = ray_vs_plane(e0.x, e0.y, e0.z, e0.w, e1.x, dirx, diry, L, val);
                            if (ok && e1.y == 0.f) v0 = val;
                            if (ok && e1.y != 0.f) v1 = val;
                        }
                        if (hz & kHdrIn0) v0 = L;       // origin inside the bank: alpha = 0, `point` stays at the ray end
                        if (hz & kHdrIn1) v1 = L;
                        // LiDAR.query: the first bank (list order) that reports a hit wins; misses keep the old reading
                        // (sticky vals, models.py:71)
                        const float v = v0 >= 0.f ? v0 : v1;
                        if (v >= 0.f) sts1(ra + 4 * (CF + 6 + rj), v);
                    }
                }
            }

            // ---- ShipGame.handle_discrete_action (game.py:140-153); Ship.move_forward / rotate (models.py:129-146)
            // The rudder angle lives in the tile, as the third value of the newest frame (exact: a small integer in fp32).
            float dvx = 0.f, dvy = 0.f, dw = 0.f;
            const unsigned rud_slot = EA(OBS4 - 4) + 8;
            const float rud = lds1(rud_slot);
            if (G > 1) __syncwarp();            // every lane has the old angle before the leader stores the new one
            if (a == 0) {                       // thrust along the heading the step starts from
                dvx = -p.acc_dt * h0.y; dvy = p.acc_dt * h0.x; dw = -p.ang_dt * rud;
            }
            else if (a == 1) { if (gl == 0) sts1(rud_slot, fmaxf(rud - 5.f, -10.f)); }
            else if (a == 2) { if (gl == 0) sts1(rud_slot, fminf(rud + 5.f, 10.f)); }
            // ---- cpBodyUpdateVelocity: v = v*damping + f/m*dt, w = w*damping + t/I*dt.  Nothing between here and the next
            // cpBodyUpdatePosition looks at the velocities (the overlap tests below work on the pose), so they are
            // advanced at once instead of carrying the three force terms across the cooperative passes.
            r.vx = r.vx * p.damping + dvx;
            r.vy = r.vy * p.damping + dvy;
            r.w = r.w * p.damping + dw;
        }
        __syncwarp();                           // the plane rows have been read (thrust heading, ray pass): they may be rewritten

        // the reach-grid cell of the integrated pose was asked for an iteration ago (asynchronous copy): it names the
        // candidate planes, whose raw records are now copied global -> shared the same way -- with G = 1 into the plane
        // row itself, which the ray pass has finished with, otherwise into the env's own staging area
        cp_async_wait_all();
        if (G > 1) __syncwarp();                // the leader's copy has landed before the group's other lanes look
        const uint4 cell = lds4u(EA(CEL));
        const bool near_any = leader && (cell.x | cell.y | (cell.z & 3u)) != 0u;
        const bool staged = near_any && ((cell.z >> 8) & 0xffu) <= (unsigned)kMaxCand;
        if (staged) {
            const float4 *E4 = reinterpret_cast<const float4 *>(p.edges_d + (size_t)r.scen * (2 * kMaxHull));
#pragma unroll
            for (int n = 0; n < kMaxCand; ++n) {
                const unsigned idx = (cell.w >> (8 * n)) & 0xffu;
                if (idx != 0xffu) {
                    cp_async16_s(EA((G > 1 ? RAW : SCR + 1) + 2 * n), E4 + 2 * idx);
                    cp_async16_s(EA((G > 1 ? RAW : SCR + 1) + 2 * n + 1), E4 + 2 * idx + 1);
                }
            }
        }
        bool all_goals = false;
        if (live) {     // (while the candidate planes are in flight)
            // ---- goals (collide_goal, game.py:243-257) and the nearest remaining goal (closest_goal, game.py:333-349).
            // The squared distance to the body origin serves both the nearest-goal search and a bounding-circle cull;
            // only goals inside the circle are rotated into the body frame for the exact circle-vs-hull test.  Taken
            // goals sit at kDeadGoal: their distance is +inf, so neither the cull nor the search needs the alive mask.
            float2 g[kGoals];
            float gd2[kGoals];
            {
                const float4 ga = lds4(EA(GOL)), gb = lds4(EA(GOL + 1));
                const float2 gc = lds2(EA(GOL + 2));
                g[0] = make_float2(ga.x, ga.y); g[1] = make_float2(ga.z, ga.w); g[2] = make_float2(gb.x, gb.y);
                g[3] = make_float2(gb.z, gb.w); g[4] = gc;
            }
            if (G > 1) __syncwarp();            // every lane has the goals before the leader retires one
            unsigned cand = 0u;
#pragma unroll
            for (int i = 0; i < kGoals; ++i) {
                const float ux = g[i].x - r.x, uy = g[i].y - r.y;
                gd2[i] = ux * ux + uy * uy;
                if (gd2[i] <= p.goal_cull_r2) cand |= 1u << i;
            }
            if (!valid) cand = 0u;
            float took = 0.f;                   // (a float: as a bool set inside the loop this flag was spilled to local memory)
#pragma unroll 1
            while (cand) {                      // rarely more than one trip
                const int i = __ffs(cand) - 1;
                cand &= cand - 1u;
                float ux = g[0].x, uy = g[0].y;
#pragma unroll
                for (int j = 1; j < kGoals; ++j) if (i == j) { ux = g[j].x; uy = g[j].y; }
                ux -= r.x; uy -= r.y;
                const float qx = ux * c + uy * s, qy = -ux * s + uy * c;
                if (goal_contact(p, qx, qy)) {
                    took = 1.f;
#pragma unroll
                    for (int j = 0; j < kGoals; ++j) if (i == j) gd2[j] = __int_as_float(0x7f800000);
                    if (gl == 0) {
                        sts2(EA(GOL) + 8 * i, make_float2(kDeadGoal, kDeadGoal));
                        sts1(EA(GOL + 2) + 8, __int_as_float(__float_as_int(lds1(EA(GOL + 2) + 8)) & ~(1 << i)));
                    }
                }
            }
            goal_reached = took != 0.f;
            float best = 3.0e38f;
#pragma unroll
            for (int i = 0; i < kGoals; ++i)
                if (gd2[i] < best) { best = gd2[i]; gx = g[i].x; gy = g[i].y; }
            all_goals = !(best < 3.0e38f);            // every goal taken: no finite distance left

        }
        // ---- plane phase at the integrated pose: next step's lidar planes + this step's ship-vs-bank pre-test
        cp_async_wait_all();
        unsigned ask = 0u;
        if (leader) {
            if (staged) ask = plane_phase<true, true>(p, r.x, r.y, hx, hy, c, s, r.scen, cell, RowSmem{EA(SCR)}, RowSmem{G > 1 ? EA(RAW) : EA(SCR + 1)});
            else if (near_any) ask = plane_phase<true, false>(p, r.x, r.y, hx, hy, c, s, r.scen, cell, RowSmem{EA(SCR)}, RowSmem{0u});
            else sts4(EA(SCR), make_float4(c, s, 0.f, 0.f));
        }

        if (live) {
            // ---- overlap test at the new pose -> begin callback collide_ship (game.py:232-241)
            bool colliding = false;
            {
                // Separating-axis test for the envs the plane phase could not settle.  Contact <=> no separating axis
                // among the edge normals of both convex polygons (touching counts: GJK distance <= 0).
                unsigned needs = __ballot_sync(kFull, ask != 0u);
                if (needs) {
                    const int lps_sh = p.hull_max <= 8 ? 3 : (p.hull_max <= 16 ? 4 : 5);    // log2(lanes per env slot)
                    const int lps = 1 << lps_sh, nslots = 32 >> lps_sh;
                    const int sslot = lane >> lps_sh, sel = lane & (lps - 1);
                    const unsigned slotmask = lps == 32 ? kFull : (((1u << lps) - 1u) << (sslot * lps));
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
                        const int srcl2 = act_env ? src : lane;
                        const float bx = __shfl_sync(kFull, r.x, srcl2), by = __shfl_sync(kFull, r.y, srcl2);
                        const float bc = __shfl_sync(kFull, c, srcl2), bs = __shfl_sync(kFull, s, srcl2);
                        const unsigned bsa = __shfl_sync(kFull, (unsigned)r.scen | (ask << 28), srcl2);
                        const float4 *rec = p.bank + (size_t)(bsa & 0x0fffffffu) * p.scen_stride4;
                        const float4 hdr = __ldg(rec + 4);
                        const float4 *bE = rec + kBankHeader4;
                        // the edge records of both banks are requested together with the header: their addresses do not
                        // depend on it (every slot below maxv exists), so one memory round trip serves the whole pass
                        float4 edb[2];
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            edb[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (act_env && ((bsa >> (28 + b)) & 1u) && sel < p.maxv) edb[b] = __ldg(bE + b * p.maxv + sel);
                        }
                        float rx[kShipVerts], ry[kShipVerts];
#pragma unroll
                        for (int j = 0; j < kShipVerts; ++j) {
                            rx[j] = p.ship_lx[j] * bc - p.ship_ly[j] * bs;
                            ry[j] = p.ship_lx[j] * bs + p.ship_ly[j] * bc;
                        }
                        bool coll = false;
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            const bool do_b = act_env && ((bsa >> (28 + b)) & 1u) && !coll;
                            const int nb = __float_as_int(b ? hdr.w : hdr.z);
                            const bool actl = do_b && sel < nb;
                            const float4 ed = edb[b];
                            const unsigned sb = __ballot_sync(kFull, actl && bank_axis_separates(ed, rx, ry, bx, by));   // a bank edge normal separates
                            bool sep = (sb & slotmask) != 0u;
                            if (__ballot_sync(kFull, do_b && !sep)) {                           // else try the ship's edge normals
#pragma unroll
                                for (int j = 0; j < kShipVerts; ++j) {
                                    const float nx = p.ship_nx[j] * bc - p.ship_ny[j] * bs;
                                    const float ny = p.ship_nx[j] * bs + p.ship_ny[j] * bc;
                                    const float pr = actl ? nx * (ed.z - bx) + ny * (ed.w - by) : 3.0e38f;
                                    // axis j separates <=> no bank vertex of the slot reaches the hull's plane j
                                    const unsigned reach = __ballot_sync(kFull, pr <= p.ship_off[j]);
                                    sep = sep || (reach & slotmask) == 0u;
                                }
                            }
                            if (do_b && !sep) coll = true;
                        }
                        const unsigned res = __ballot_sync(kFull, coll);
                        if (myslot >= 0 && ((res >> (myslot * lps)) & 1u)) colliding = true;
                    }
                }
            }

            // ---- ShipEnv.determine_reward (ship_env.py:62-77): collision alone does not change the value (Q12)
            const bool oob = (r.x < 0.f) || (r.x > p.W) || (r.y < 0.f) || (r.y > p.H);
            reward = goal_reached ? 1.f : (oob ? -1.f : p.step_penalty);
            r.ret += reward;
            r.steps += 1;
            const bool timeout = (r.steps >= p.max_steps);
            done = colliding || all_goals || oob || timeout;             // ship_env.py:115-134

            warp_stats(wb + STA * 16, lane, leader, goal_reached, done, colliding, oob, timeout, all_goals, r.ret, r.steps);
            do_reset = done && p.auto_reset;
            int ep_g = 0;
            if (G > 1) {                        // read by every lane while the warp is converged, before the leader stores the next one
                ep_g = __float_as_int(lds1(EA(GOL + 2) + 12)) + 1;
                __syncwarp();
            }
            if (do_reset) {
                const int ep = G > 1 ? ep_g : __float_as_int(lds1(EA(GOL + 2) + 12)) + 1;    // (only a reset needs the episode number: kept in shared memory)
                reset_env(p, r, pick_scenario(p, p.env_id_offset + e, ep), ep);
                c = 1.f; s = 0.f;
                hx = 0.5f * (p.ship_aabb[2] - p.ship_aabb[0]); hy = 0.5f * (p.ship_aabb[3] - p.ship_aabb[1]);
                {
                    const float4 *rec = p.bank + (size_t)r.scen * p.scen_stride4;
                    float2 gn[kGoals];
                    const float4 rg0 = __ldg(rec + 2), rg1 = __ldg(rec + 3), rg2 = __ldg(rec + 4);
                    unpack_goals(rg0, rg1, rg2, gn);
                    closest_goal(gn, r.alive, r.x, r.y, gx, gy);
                    if (gl == 0) {
                        sts4(EA(GOL), rg0); sts4(EA(GOL + 1), rg1);
                        sts4(EA(GOL + 2), make_float4(rg2.x, rg2.y, __int_as_float((1 << kGoals) - 1), __int_as_float(ep)));
                    }
                    // the goal planes of the state change only here: written at once (a "goals changed" flag carried
                    // to the end of the kernel had been spilled to local memory and reloaded every iteration)
                    if (leader) store_goals(p, e, rg0, rg1, rg2);
                }
                if (gl == 0) {              // the spawn pose's planes were evaluated when the scenario was loaded
                    const float4 *sp = p.spawn_rows + (size_t)r.scen * kScr4;
                    const float4 h0 = __ldg(sp);
                    const int hn = __float_as_int(h0.z);
                    const int nrow = (hn & kHdrBig) ? 2 : 2 * (hn & 0xff);
                    sts4(EA(SCR), h0);
                    for (int i = 1; i <= nrow; ++i) sts4(EA(SCR + i), __ldg(sp + i));
                }
            }
        }
        if (live && gl == 0) {
            // ---- this step's observation frame.  The leader completes the newest frame in the resident tile (the lidar
            // slots are already there) BEFORE the pose moves on, so that no copy of this step's pose has to be kept.
            if (do_reset) {                                     // ship_env.py:180-184: [-1 x 16 | reset frame], vals = -1
                const float4 neg = make_float4(-1.f, -1.f, -1.f, -1.f);
                if (HIST == 2) { sts4(EA(0), neg); sts4(EA(1), neg); sts4(EA(2), neg); sts4(EA(3), neg); }
                sts4(EA(OBS4 - 4), make_float4(r.x, r.y, 0.f, r.th));       // rudder 0 (models.py:108)
                sts4(EA(OBS4 - 3), make_float4(gx, gy, -1.f, -1.f));
                sts4(EA(OBS4 - 2), neg);
                sts4(EA(OBS4 - 1), neg);
            } else {                            // the rudder slot already holds this step's angle
                sts2(EA(OBS4 - 4), make_float2(r.x, r.y));
                sts1(EA(OBS4 - 4) + 12, r.th);
                sts2(EA(OBS4 - 3), make_float2(gx, gy));
            }
        }
        // ---- cpSpaceStep of the NEXT step, positions first (cpBodyUpdatePosition).  Done before this step's copy-out
        // so that the grid cell of the new pose is requested as early as possible: it is consumed an iteration later.
        if (k + 1 < p.K) {
            r.x += r.vx * p.dt;
            r.y += r.vy * p.dt;
            r.th += r.w * p.dt;
            sincos_fast(r.th, s, c);
            hull_half_extents(p, c, s, hx, hy);
            if (G == 1 || gl == 0) fetch_cell_async(p, r.scen, r.x + hx, r.y + hy, EA(CEL));
        }
        if (live) {
            // ---- outputs: obs rows of the warp's envs are contiguous in global memory, so the tile is copied out with
            // fully coalesced 128-bit streaming stores.
            if (SHIPSIM_TMA_STORE) fence_proxy_async();     // this lane's writes to the tiles (its own frame, other envs' rays)
            __syncwarp();
#if SHIPSIM_TMA_STORE
            if (p.obs && leader) bulk_store(p.obs + ((size_t)k * p.N + e) * OBS4, EA(0), OBS4 * 16);
#else
            // lane -> (row lane / OBS4, column lane % OBS4), each further round moves 32 / OBS4 rows down.  A round is real
            // iff its row belongs to an env of this batch, i.e. iff its address lies before the end of this step's rows
            // (tested against the step's end pointer, which moves with k: a loop-invariant row limit was hoisted, spilled
            // and reloaded from local memory right here in every iteration)
            if (p.obs && lane < EPW * OBS4) {
                const unsigned cp_src = WA(lane / OBS4, lane % OBS4);
                float4 *o = p.obs + ((size_t)k * p.N + (e - grp)) * OBS4 + lane;
                const float4 *o_end = p.obs + (size_t)(k + 1) * p.N * OBS4;
#pragma unroll
                for (int i = 0; i < (EPW * OBS4 >= 32 ? EPW * OBS4 / 32 : 1); ++i)
                    if (o + i * 32 < o_end) __stcs(o + i * 32, lds4(cp_src + i * (32 / OBS4) * (EB4 * 16)));
            }
#endif
            if (leader) {
                const size_t row = (size_t)k * p.N + e;
                if (p.reward) p.reward[row] = reward;
                if (p.done) p.done[row] = done ? 1 : 0;
            }
        }
        __syncwarp();                           // copy-out done and plane rows complete before the next iteration
    }
    // the loop leaves the pose of the last step in r (no integration after it)
    if (SHIPSIM_TMA_STORE) bulk_store_wait_read();
    if (leader) {
        const float4 l1 = lds4(EA(OBS4 - 3)), l2 = lds4(EA(OBS4 - 2)), l3 = lds4(EA(OBS4 - 1));
        r.episode = __float_as_int(lds1(EA(GOL + 2) + 12));
        r.alive = __float_as_int(lds1(EA(GOL + 2) + 8));
        r.rudder = (int)lds1(EA(OBS4 - 4) + 8);
        store_env(p, e, r, make_float4(l1.z, l1.w, l2.x, l2.y), make_float4(l2.z, l2.w, l3.x, l3.y), l3.z, l3.w);
    }

    // episode statistics: one red.add per non-zero value per warp into a slot row
    __syncwarp();
    if (lane < 8 && p.stats) {
        const float v = lds1(wb + STA * 16 + 4 * lane);
        if (v != 0.f) atomicAdd(p.stats + (size_t)(blockIdx.x % kStatSlots) * kStatLen + lane, (double)v);
    }
#undef EA
#undef WA
}

// plane phase at the spawn pose of every scenario (what a reset env starts from), one thread per scenario
__global__ void __launch_bounds__(128) build_spawn_rows_kernel(const __grid_constant__ StepParams p, float4 *rows)
{
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= p.n_scen) return;
    float4 *row = rows + (size_t)sidx * kScr4;
    for (int i = 0; i < kScr4; ++i) row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float hx = 0.5f * (p.ship_aabb[2] - p.ship_aabb[0]), hy = 0.5f * (p.ship_aabb[3] - p.ship_aabb[1]);
    const uint4 cell = load_cell(p, sidx, p.spawn_x + hx, p.spawn_y + hy);
    if ((cell.x | cell.y | (cell.z & 3u)) != 0u) plane_phase<false, false>(p, p.spawn_x, p.spawn_y, hx, hy, 1.f, 0.f, sidx, cell, RowPtr{row}, RowPtr{nullptr});
    else row[0] = make_float4(1.f, 0.f, 0.f, 0.f);
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
    // w: the first (up to) four candidates as bytes, bank * kMaxHull + edge, in the order the plane phase evaluates them
    // (bank 0 first, ascending edge index); 0xff = none.  z bits 8..15: the number of candidates.
    unsigned list = 0xffffffffu;
    int cnt = 0;
    for (int b = 0; b < 2; ++b)
        for (unsigned m = masks[b]; m; m &= m - 1u) {
            if (cnt < 4) list = (list & ~(0xffu << (8 * cnt))) | ((unsigned)(b * kMaxHull + (__ffs(m) - 1)) << (8 * cnt));
            ++cnt;
        }
    grid[t] = make_uint4(masks[0], masks[1], flags | ((unsigned)cnt << 8), list);
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
    float4 l0, l1, l2, g0, g1, g2;
    load_env(p, e, r, l0, l1, l2, g0, g1, g2);
    const int ep = first ? 0 : r.episode + 1;
    int scen = scenario ? scenario[e] : pick_scenario(p, p.env_id_offset + e, ep);
    scen = min(max(scen, 0), p.n_scen - 1);       // caller-supplied ids index the bank: never out of range (the host layer raises first)
    reset_env(p, r, scen, ep);
    const float4 *rec = p.bank + (size_t)scen * p.scen_stride4;
    g0 = __ldg(rec + 2); g1 = __ldg(rec + 3); g2 = __ldg(rec + 4);
    const float4 neg4 = make_float4(-1.f, -1.f, -1.f, -1.f);      // models.py:36: LiDAR.vals start at -1
    store_env(p, e, r, neg4, neg4, -1.f, -1.f);
    store_goals(p, e, g0, g1, g2);
    if (obs) {
        float gx, gy;
        float2 g[kGoals];
        unpack_goals(g0, g1, g2, g);
        closest_goal(g, r.alive, r.x, r.y, gx, gy);
        float4 *o = obs + (size_t)e * (4 * p.history);
        const float4 neg = make_float4(-1.f, -1.f, -1.f, -1.f);
        if (p.history == 2) { o[0] = neg; o[1] = neg; o[2] = neg; o[3] = neg; o += 4; }
        o[0] = make_float4(r.x, r.y, 0.f, 0.f);
        o[1] = make_float4(gx, gy, -1.f, -1.f);
        o[2] = neg;
        o[3] = neg;
    }
}

// After a bank swap (shipsim_load_scenarios / shipsim_generate_scenarios on a live handle) a stored scenario id may
// exceed the new bank: fold it into range so that no kernel indexes bank / grid / edges_d / spawn_rows out of bounds.
// (The env keeps its old goals until its next reset; the host layer resets the whole batch after a swap.)
__global__ void __launch_bounds__(256) clamp_scenarios_kernel(float4 *state, int N, int n_scen)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    float4 *q = state + (size_t)4 * N + e;
    const int scen = __float_as_int(q->z);
    if (scen < 0 || scen >= n_scen) q->z = __int_as_float((int)((unsigned)scen % (unsigned)n_scen));
}

// the observation frame of the CURRENT state of every env (ShipEnv.__add_states, ship_env.py:79-113): what the next
// step will report as its "previous" frame.  Used by shipsim_step_host, which ships frames and rebuilds the history.
__global__ void __launch_bounds__(256) frame_kernel(const __grid_constant__ StepParams p, float4 *out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.N) return;
    EnvRegs r;
    float4 l0, l1, l2, g0, g1, g2;
    load_env(p, e, r, l0, l1, l2, g0, g1, g2);
    float2 g[kGoals];
    unpack_goals(g0, g1, g2, g);
    float gx, gy;
    closest_goal(g, r.alive, r.x, r.y, gx, gy);
    float4 *o = out + (size_t)e * 4;
    o[0] = make_float4(r.x, r.y, (float)r.rudder, r.th);
    o[1] = make_float4(gx, gy, l0.x, l0.y);
    o[2] = make_float4(l0.z, l0.w, l1.x, l1.y);
    o[3] = make_float4(l1.z, l1.w, l2.x, l2.y);
}

// ------------------------------------------------------------------------------------------------------------
// Lossless compaction of a chunk of frames for the host path (shipsim_step_host): 64 + 5 bytes per env-step shrink to
// ~20, which is what crosses PCIe.  Between consecutive frames of an env the pose changes every step, everything else
// rarely: rudder / reward / done are a handful of states, the nearest goal switches now and then, a lidar reading only
// when its ray hits.  Per env-step the kernel emits one 16-byte record
//     { x, y, angle, word }    word = rudder code | reward code << 3 | done << 5 | change mask << 8
// (change mask: bit j set <=> slot 4 + j of the frame -- gx, gy, lidar 0..9 -- differs bitwise from the env's previous
// frame) and appends the changed values to a variable-length stream; off[k][block] = where the values of the 32 envs
// of a block start at step k (the blocks' order in the stream is whatever the atomics made it: the table is the truth).
// One warp per (step, 32-env block), blocks blk0 and up (the envs below 32 * blk0 go home as complete rows by DMA:
// history_rows_kernel).  The host rebuilds the frames exactly (shipsim_host.cpp: expand_delta_rows).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_frames_kernel(const float4 *frames, const float4 *prev0, const float *rew, const uint8_t *done,
                                                             int N, int blk0, int kc, float step_penalty, uint4 *rec, unsigned *off, float *var,
                                                             unsigned var_cap, unsigned *counter)
{
    const int lane = threadIdx.x & 31;
    const int nblk = (N + 31) / 32, nb = nblk - blk0;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)kc * nb) return;
    const int k = (int)(w / nb), blk = blk0 + (int)(w % nb);
    const int e = blk * 32 + lane;
    const bool valid = e < N;
    unsigned mask = 0u;
    float vals[12];
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
        const float4 *f = frames + ((size_t)k * N + e) * 4;
        const float4 *q = k > 0 ? f - (size_t)N * 4 : prev0 + (size_t)e * 4;
        const float4 f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3];
        const float4 q1 = q[1], q2 = q[2], q3 = q[3];
        const float nv[12] = {f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w, f3.x, f3.y, f3.z, f3.w};
        const float ov[12] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            vals[j] = nv[j];
            if (__float_as_uint(nv[j]) != __float_as_uint(ov[j])) mask |= 1u << j;
        }
        const float rw = rew[(size_t)k * N + e];
        const unsigned rcode = rw == 1.f ? 1u : (rw == -1.f ? 2u : (rw == step_penalty ? 0u : 3u));     // 3 never happens
        const unsigned rud = (unsigned)((int)f0.z / 5 + 2) & 7u;
        r = make_uint4(__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.w),
                       rud | (rcode << 3) | ((done[(size_t)k * N + e] ? 1u : 0u) << 5) | (mask << 8));
        rec[(size_t)k * N + e] = r;
    }
    const int n = __popc(mask);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    unsigned base = 0u;
    if (lane == 0) {
        base = total ? atomicAdd(counter, (unsigned)total) : 0u;
        off[(size_t)k * nblk + blk] = base;
    }
    base = __shfl_sync(kFull, base, 0);
    unsigned at = base + (unsigned)(incl - n);
#pragma unroll
    for (int j = 0; j < 12; ++j)
        if ((mask >> j) & 1u) {
            if (at < var_cap) var[at] = vals[j];        // (beyond the capacity: dropped; the host sees counter > capacity)
            ++at;
        }
}

// Complete observation rows [previous frame | frame] of the envs [0, nd) for a chunk of frames (ship_env.py:112-113;
// 16 x -1 as the previous frame of a step that ended an episode under auto-reset, ship_env.py:180-184): the part of the
// host path's output that travels as plain rows by DMA while the host threads expand the compacted part.  One thread per
// quarter frame; rows[kc][nd][32].
__global__ void __launch_bounds__(256) history_rows_kernel(const float4 *frames, const float4 *prev0, const uint8_t *done, int N, int nd,
                                                           int kc, int cut, float4 *rows)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)kc * nd * 4) return;
    const int q = (int)(t & 3);
    const long long ke = t >> 2;
    const int k = (int)(ke / nd), e = (int)(ke % nd);
    const size_t at = ((size_t)k * N + e) * 4 + q;
    const float4 cur = frames[at];
    float4 prev = k > 0 ? frames[at - (size_t)N * 4] : prev0[(size_t)e * 4 + q];
    if (cut && done[(size_t)k * N + e]) prev = make_float4(-1.f, -1.f, -1.f, -1.f);
    float4 *o = rows + (size_t)ke * 8 + q;
    __stcs(o, prev);
    __stcs(o + 4, cur);
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
    int blocks = (p.N + envs_per_cta - 1) / envs_per_cta;
    int threads = kThreads;
    bool launched = false;
    if constexpr (G == 1) {
        if (blocks >= 148 * 8) {
            threads = SHIPSIM_BIG_THREADS;
            blocks = (p.N + threads - 1) / threads;
            if (p.history == 2) step_kernel<G, 2, SHIPSIM_BIG_MIN_BLOCKS, SHIPSIM_BIG_THREADS><<<blocks, threads, 0, stream>>>(p);
            else step_kernel<G, 1, SHIPSIM_BIG_MIN_BLOCKS, SHIPSIM_BIG_THREADS><<<blocks, threads, 0, stream>>>(p);
            launched = true;
        }
    }
    if (!launched) {
        // Grids that do not fill the machine: 64-thread CTAs (same 16 warps per SM at 128 registers), so that the few
        // CTAs spread evenly -- 65,536 envs are 512 CTAs of 128 threads, 3.46 per SM: every SM waits for the ones that got 4
        constexpr int T = SHIPSIM_SMALL_THREADS, MB = SHIPSIM_MIN_BLOCKS * kThreads / SHIPSIM_SMALL_THREADS;
        threads = T;
        blocks = (p.N + T / G - 1) / (T / G);
        if (p.history == 2) step_kernel<G, 2, MB, T><<<blocks, T, 0, stream>>>(p);
        else step_kernel<G, 1, MB, T><<<blocks, T, 0, stream>>>(p);
    }
    if (shape) { shape->lanes_per_env = G; shape->threads = threads; shape->blocks = blocks; shape->window = 1; }
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

cudaError_t launch_clamp_scenarios(float4 *state, int N, int n_scen, cudaStream_t stream)
{
    clamp_scenarios_kernel<<<(N + 255) / 256, 256, 0, stream>>>(state, N, n_scen);
    return cudaGetLastError();
}

cudaError_t launch_compact_frames(const float4 *frames, const float4 *prev0, const float *rew, const uint8_t *done, int N, int env0, int kc,
                                  float step_penalty, uint4 *rec, unsigned *off, float *var, unsigned var_cap, unsigned *counter,
                                  cudaStream_t stream)
{
    const long long warps = (long long)kc * ((N + 31) / 32 - env0 / 32);
    if (warps <= 0) return cudaSuccess;
    compact_frames_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(frames, prev0, rew, done, N, env0 / 32, kc, step_penalty,
                                                                                   rec, off, var, var_cap, counter);
    return cudaGetLastError();
}

cudaError_t launch_history_rows(const float4 *frames, const float4 *prev0, const uint8_t *done, int N, int nd, int kc, int cut, float4 *rows,
                                cudaStream_t stream)
{
    const long long threads = (long long)kc * nd * 4;
    if (threads <= 0) return cudaSuccess;
    history_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(frames, prev0, done, N, nd, kc, cut, rows);
    return cudaGetLastError();
}

cudaError_t launch_frame(const StepParams &p, float4 *out, cudaStream_t stream)
{
    frame_kernel<<<(p.N + 255) / 256, 256, 0, stream>>>(p, out);
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

cudaError_t launch_build_spawn_rows(const StepParams &p, float4 *rows, cudaStream_t stream)
{
    build_spawn_rows_kernel<<<(p.n_scen + 127) / 128, 128, 0, stream>>>(p, rows);
    return cudaGetLastError();
}

cudaError_t launch_stats_reduce(double *slots, double *out, int clear, cudaStream_t stream)
{
    stats_reduce_kernel<<<1, 32 * kStatLen, 0, stream>>>(slots, out, clear);
    return cudaGetLastError();
}

}  // namespace shipsim
