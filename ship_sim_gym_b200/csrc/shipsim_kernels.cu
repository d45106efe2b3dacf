// shipsim_kernels.cu -- the fused ShipEnv step kernel (K env-steps per launch), reset and stats kernels.
// sm_100a only.  See shipsim_device.cuh for the data layout and the reference lines each piece restates.
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

__device__ __forceinline__ void st_stream(float4 *ptr, float4 v) { __stcs(ptr, v); }

// ------------------------------------------------------------------------------------------------------------
// One thread per env.  State is loaded once, lives in registers for K steps, and is stored once.
// ------------------------------------------------------------------------------------------------------------
template <int HIST>
__global__ void __launch_bounds__(kThreadsT1) step_kernel_t1(const __grid_constant__ StepParams p)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    float st_episodes = 0.f, st_return = 0.f, st_length = 0.f, st_goal = 0.f;
    float st_coll = 0.f, st_oob = 0.f, st_timeout = 0.f, st_allgoals = 0.f;

    if (e < p.N) {
        EnvRegs r;
        load_env(p, e, r);
        const long long gid = p.env_id_offset + e;
        float c, s;
        sincosf(r.th, &s, &c);
        float gx, gy;
        closest_goal(r, gx, gy);
        bool goals_dirty = false;

#pragma unroll 1
        for (int k = 0; k < p.K; ++k) {
            const int a = load_action(p, k, e, gid);
            // previous frame = newest frame of the last step / reset (SURVEY.md App. A note N2)
            float4 P0 = make_float4(r.x, r.y, (float)r.rudder, r.th);
            float4 P1 = make_float4(gx, gy, r.lid[0], r.lid[1]);
            float4 P2 = make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]);
            float4 P3 = make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]);
            const float4 *sc = p.bank + (size_t)r.scen * p.scen_stride4;

            // ShipGame.handle_discrete_action (game.py:140-153); Ship.move_forward / rotate (models.py:129-146)
            float dvx = 0.f, dvy = 0.f, dw = 0.f;
            if (a == 0) { dvx = -p.acc_dt * s; dvy = p.acc_dt * c; dw = -p.ang_dt * (float)r.rudder; }
            else if (a == 1) r.rudder = max(r.rudder - 5, -10);
            else if (a == 2) r.rudder = min(r.rudder + 5, 10);

            lidar_query(p, sc, r, c, s);                       // game.py:193

            // cpSpaceStep: positions first (cpBodyUpdatePosition)
            r.x += r.vx * p.dt;
            r.y += r.vy * p.dt;
            r.th += r.w * p.dt;
            sincosf(r.th, &s, &c);

            // overlap tests at the new pose -> begin callbacks collide_ship / collide_goal (game.py:232-257)
            float rx[kShipVerts], ry[kShipVerts];
            float sminx = 3.0e38f, sminy = 3.0e38f, smaxx = -3.0e38f, smaxy = -3.0e38f;
#pragma unroll
            for (int j = 0; j < kShipVerts; ++j) {
                rx[j] = p.ship_lx[j] * c - p.ship_ly[j] * s;
                ry[j] = p.ship_lx[j] * s + p.ship_ly[j] * c;
                sminx = fminf(sminx, r.x + rx[j]); smaxx = fmaxf(smaxx, r.x + rx[j]);
                sminy = fminf(sminy, r.y + ry[j]); smaxy = fmaxf(smaxy, r.y + ry[j]);
            }
            const float4 hdr = __ldg(sc + 4);
            const bool colliding =
                ship_touches_bank(p, sc, 0, __float_as_int(hdr.z), r.x, r.y, rx, ry, c, s, sminx, sminy, smaxx, smaxy) ||
                ship_touches_bank(p, sc, 1, __float_as_int(hdr.w), r.x, r.y, rx, ry, c, s, sminx, sminy, smaxx, smaxy);
            bool goal_reached = false;
#pragma unroll
            for (int g = 0; g < kGoals; ++g) {
                if ((r.alive >> g) & 1) {
                    const float rx = r.g[2 * g] - r.x, ry = r.g[2 * g + 1] - r.y;
                    if (goal_touches_ship(p, rx * c + ry * s, -rx * s + ry * c)) {
                        goal_reached = true;
                        r.alive &= ~(1 << g);
                    }
                }
            }

            // cpBodyUpdateVelocity: v = v*damping + f/m*dt, w = w*damping + t/I*dt
            r.vx = r.vx * p.damping + dvx;
            r.vy = r.vy * p.damping + dvy;
            r.w = r.w * p.damping + dw;

            // ShipEnv.determine_reward (ship_env.py:62-77): collision alone does not change the value (Q12)
            const bool oob = (r.x < 0.f) || (r.x > p.W) || (r.y < 0.f) || (r.y > p.H);
            const float reward = goal_reached ? 1.f : (oob ? -1.f : p.step_penalty);
            r.ret += reward;
            r.steps += 1;
            closest_goal(r, gx, gy);
            const bool all_goals = (r.alive == 0);
            const bool timeout = (r.steps >= p.max_steps);
            const bool done = colliding || all_goals || oob || timeout;      // ship_env.py:115-134

            st_goal += goal_reached ? 1.f : 0.f;
            if (done) {
                st_episodes += 1.f; st_return += r.ret; st_length += (float)r.steps;
                st_coll += colliding ? 1.f : 0.f; st_oob += oob ? 1.f : 0.f;
                st_timeout += timeout ? 1.f : 0.f; st_allgoals += all_goals ? 1.f : 0.f;
                if (p.auto_reset) {
                    const int ep = r.episode + 1;
                    reset_env(p, r, pick_scenario(p, gid, ep), ep);
                    c = 1.f; s = 0.f;
                    closest_goal(r, gx, gy);
                    goals_dirty = true;
                    P0 = P1 = P2 = P3 = make_float4(-1.f, -1.f, -1.f, -1.f);   // ship_env.py:180-181
                }
            }

            const size_t row = (size_t)k * p.N + e;
            if (p.obs) {
                float4 *o = p.obs + row * (4 * HIST);
                if (HIST == 2) { st_stream(o + 0, P0); st_stream(o + 1, P1); st_stream(o + 2, P2); st_stream(o + 3, P3); o += 4; }
                st_stream(o + 0, make_float4(r.x, r.y, (float)r.rudder, r.th));
                st_stream(o + 1, make_float4(gx, gy, r.lid[0], r.lid[1]));
                st_stream(o + 2, make_float4(r.lid[2], r.lid[3], r.lid[4], r.lid[5]));
                st_stream(o + 3, make_float4(r.lid[6], r.lid[7], r.lid[8], r.lid[9]));
            }
            if (p.reward) p.reward[row] = reward;
            if (p.done) p.done[row] = done ? 1 : 0;
        }
        store_env(p, e, r, goals_dirty);
    }

    // episode statistics: warp shuffle reduction, then one red.add per non-zero value per warp into a slot row
    float v[8] = {st_episodes, st_return, st_length, st_goal, st_coll, st_oob, st_timeout, st_allgoals};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
    if ((threadIdx.x & 31) == 0 && p.stats) {
        double *row = p.stats + (size_t)(blockIdx.x % kStatSlots) * kStatLen;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (v[i] != 0.f) atomicAdd(row + i, (double)v[i]);
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
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[col] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// launch wrappers (called from the C ABI)
// ------------------------------------------------------------------------------------------------------------
cudaError_t launch_step(const StepParams &p, int lanes_per_env, cudaStream_t stream, LaunchShape *shape)
{
    (void)lanes_per_env;
    const int threads = kThreadsT1;
    const int blocks = (p.N + threads - 1) / threads;
    if (shape) { shape->lanes_per_env = 1; shape->threads = threads; shape->blocks = blocks; }
    if (p.history == 2) step_kernel_t1<2><<<blocks, threads, 0, stream>>>(p);
    else step_kernel_t1<1><<<blocks, threads, 0, stream>>>(p);
    return cudaGetLastError();
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
