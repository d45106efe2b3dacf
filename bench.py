#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused ShipEnv step on B200, with roofline, CPU baseline and e2e numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One bench "step" = one pass of the hot path over one batch: a rollout of ROLLOUT (1000) consecutive env-steps of
ENVS (4096) environments in ONE persistent kernel launch, with synthetic uniform{0,1,2} actions resident in HBM
(BASELINE.json configs[1]: "default map, 4,096 batched envs on 1xB200, random actions, 1,000 steps").
N > 1 (torchrun, one rank per GPU): every rank runs that same workload on its own shard of global env ids
(weak scaling) plus ONE all-reduce of the 16-double episode-statistics vector per rollout -- the only collective
the path has.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's version banner), so file descriptor 1 is
# pointed at stderr for the duration of the run and the line goes to the real stdout at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


if os.environ.get("SHIPSIM_BENCH_WATCHDOG"):        # debugging aid: dump every thread's stack if the run is still going after N seconds
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ["SHIPSIM_BENCH_WATCHDOG"]), exit=True)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ENVS = 4096          # BASELINE.json configs[1]
ROLLOUT = 1000       # env-steps per launch ("1,000 steps")
N_SCENARIOS = 1024
SEED = 0
METRIC = "env-steps/sec"
UNIT = "env-steps/s"


def b_alg(K):
    """Algorithmic HBM bytes per env-step for a K-step fused launch: mandatory I/O
    action 4 + obs 128 + reward 4 + done 1 = 137, plus state read 116 + write 72 = 188 once per launch."""
    return 137.0 + 188.0 / K


def kernel_name(info):
    """The kernel the last launch ran: the time-parallel window kernel (small batches) or the serial-in-time one."""
    if info.get("steps_in_flight", 1) > 1:
        return "window_kernel<T=%d>" % info["steps_in_flight"]
    return "step_kernel<G=%d>" % info["lanes_per_env"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_port_rate(n_envs, budget_s, threads):
    """Time the float64 C restatement of the reference (oracle/, kind "port") on the host cores: same workload
    (default map, random Philox actions, auto-reset), a bounded number of env-steps."""
    import oracle
    from ship_sim_gym_b200 import ScenarioBank
    bank = ScenarioBank.generate(N_SCENARIOS, (600, 600), seed=SEED).as_dict()
    env = oracle.OracleEnv(n_envs, bank, auto_reset=True, seed=SEED, n_threads=threads)
    env.reset()
    want = ("obs", "reward", "done")
    env.step(None, K=4, want=want)                      # warm-up + calibration
    t0 = time.perf_counter()
    env.step(None, K=8, want=want)
    per_step = (time.perf_counter() - t0) / 8
    K = max(8, min(4000, int(budget_s / max(per_step, 1e-9))))
    t0 = time.perf_counter()
    env.step(None, K=K, want=want)
    dt = time.perf_counter() - t0
    return n_envs * K / dt, K, dt


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Python over
    pymunk/Chipmunk, which cannot be installed in this image, so this times the C restatement (kind "port"),
    with every host thread, each step being a bounded sample of the workload."""
    if rank != 0:
        return
    import oracle
    from ship_sim_gym_b200 import ScenarioBank
    cores = os.cpu_count() or 1
    bank = ScenarioBank.generate(N_SCENARIOS, (600, 600), seed=SEED).as_dict()
    env = oracle.OracleEnv(ENVS, bank, auto_reset=True, seed=SEED, n_threads=cores)
    env.reset()
    want = ("obs", "reward", "done")
    t0 = time.perf_counter()
    env.step(None, K=4, want=want)
    per_env_step = (time.perf_counter() - t0) / 4
    total = max(1, args.steps + args.warmup)
    sample_K = max(1, min(ROLLOUT, int(60.0 / total / max(per_env_step, 1e-9))))
    for _ in range(args.warmup):
        env.step(None, K=sample_K, want=want)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env.step(None, K=sample_K, want=want)
    dt = time.perf_counter() - t0
    value = ENVS * sample_K * args.steps / dt
    sample = "%d envs x %d env-steps per bench step (of %d), auto-reset, Philox random actions" % (ENVS, sample_K, ROLLOUT)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = pure Python over pymunk/Chipmunk (not installable here); timed arm is the float64 C "
                "restatement in oracle/ (faster than the real Python env: no cffi, no pygame, no clock.tick sleep)",
    }
    emit(line)


def workload_config(world):
    return {"workload": "default map 600x600, SPEED=10, %d envs/GPU, random actions, %d env-steps per launch "
                        "(BASELINE configs[1])" % (ENVS, ROLLOUT),
            "envs_per_gpu": ENVS, "global_envs": ENVS * world, "rollout_steps": ROLLOUT, "n_scenarios": N_SCENARIOS,
            "history": 2, "auto_reset": True, "parallelism": "env-shard x%d" % world,
            "l2": "rollout buffers (%.0f MB per launch) exceed the 126 MB L2; no explicit flush"
                  % ((ENVS * ROLLOUT * 137) / 1e6)}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra (non-headline) configurations")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--window", type=int, default=0, help="steps_in_flight (0 = auto)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 0)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank, dist as sdist

    rank, world, local = sdist.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    bank = ScenarioBank.generate(N_SCENARIOS, (600, 600), seed=SEED)
    env = BatchedShipEnv(ENVS, bank=bank, seed=SEED, auto_reset=True, device=dev, env_id_offset=rank * ENVS,
                         lanes_per_env=args.lanes, steps_in_flight=args.window, validate_actions=False)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(SEED + rank)
    actions = torch.randint(0, 3, (ROLLOUT, ENVS), dtype=torch.int32, device=dev, generator=gen)
    out = env.alloc_rollout(ROLLOUT)
    reducer = sdist.StatsReducer(dev)

    def one_step():
        env.rollout(actions, out=out)
        if world > 1:       # one all-reduce of the episode statistics per rollout; it overlaps the next rollout's kernel
            mode = os.environ.get("SHIPSIM_BENCH_REDUCE", "full")
            if mode == "full":
                env.stats_tensor(clear=True, out=reducer.next_buffer())
                reducer.reduce()
            elif mode == "local":
                env.stats_tensor(clear=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = env.launch_info()["launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_step()
    reducer.wait_all()                      # the last all-reduce belongs to the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_info()["launches"] - l0
    if sampler.nv is not None and len(sampler.samples) < 5:       # region too short for NVML: extend, untimed
        t_end = time.time() + 0.5
        while time.time() < t_end:          # no collective in here: ranks run different numbers of these
            env.rollout(actions, out=out)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    env_steps = float(ENVS) * ROLLOUT * args.steps * world
    value = env_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the fused step kernel is the only kernel in the timed region)
    peak, peak_src = load_peaks()
    launch_ms = ms / args.steps
    achieved = b_alg(ROLLOUT) * ENVS * ROLLOUT / (launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": kernel_name(env.launch_info()), "bytes_per_env_step": b_alg(ROLLOUT),
                "launch_ms": launch_ms}

    # ---- end to end through the C ABI with HOST buffers (copies inside the timed region)
    e2e = None
    pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype).pin_memory()   # noqa: E731
    h_act = pin(ROLLOUT, ENVS, dtype=torch.int32)
    h_act.copy_(actions.cpu())
    h_obs, h_rew, h_done = pin(ROLLOUT, ENVS, 32, dtype=torch.float32), pin(ROLLOUT, ENVS, dtype=torch.float32), pin(ROLLOUT, ENVS, dtype=torch.uint8)
    host_out = (h_obs.numpy(), h_rew.numpy(), h_done.numpy())
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        env.step_host(h_act.numpy(), K=ROLLOUT, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.step_host(h_act.numpy(), K=ROLLOUT, out=host_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d_b, d2h_b = env.host_traffic()
    e2e = {"value": float(ENVS) * ROLLOUT * e2e_steps * world / float(t.item()), "unit": UNIT,
           # counted by shipsim_step_host from the copies it issues.  HISTORY_SIZE = 2: only frames cross PCIe (64 B per
           # env-step) and the 128-byte [previous | current] rows are rebuilt in the caller's buffer by host threads
           "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
           "host_obs_bytes_per_step": int(h_obs.numel() * 4),
           "steps": e2e_steps, "api": "BatchedShipEnv.step_host -> shipsim_step_host (pinned host buffers)"}

    # ---- extra configurations (not the headline; same kernel)
    extra = {}
    if not args.no_extra:
        extra = extra_configs(torch, dev, rank, world)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v_all, K_all, dt_all = cpu_port_rate(ENVS, 10.0, cores)
            v_one, K_one, dt_one = cpu_port_rate(ENVS, 5.0, 1)
            cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d envs x %d env-steps (%.1f s), all host threads; single thread: %.3g env-steps/s "
                             "(%d env-steps, %.1f s)" % (ENVS, K_all, dt_all, v_one, K_one, dt_one),
                   "single_thread_value": v_one}
        info = env.launch_info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
            "launch_shape": {k: info[k] for k in ("lanes_per_env", "threads_per_cta", "ctas", "steps_in_flight")},
            "stats": env.stats(), "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_configs(torch, dev, rank, world):
    """Other BASELINE.json configurations, timed with the same kernel (reported under "extra", never as the
    headline): configs[2] hard map 65,536 envs; configs[3] 1,048,576 envs sharded over the ranks."""
    import torch.distributed as dist
    from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank, dist as sdist
    from ship_sim_gym_b200.config import EnvConfig, GameConfig
    peak, _ = load_peaks()
    res = {}

    def timed(env, K, reps, with_allreduce):
        acts = torch.randint(0, 3, (K, env.num_envs), dtype=torch.int32, device=dev)
        out = env.alloc_rollout(K)
        stats = torch.zeros(16, dtype=torch.float64, device=dev)

        def go():
            env.rollout(acts, out=out)
            if with_allreduce and world > 1:
                stats.copy_(env.stats_tensor(clear=True))
                sdist.all_reduce_stats(stats)
        for _ in range(3):
            go()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            go()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / reps
        return ms

    # configs[2]: builder-defined "max difficulty" map: 1000x1000, N=30, width_frac=0.9, 180 deg fan
    if rank == 0:
        class GC(GameConfig):
            BOUNDS = (1000, 1000)
        bank = ScenarioBank.generate(256, (1000, 1000), seed=SEED, map_N=30, width_frac=0.9)
        env = BatchedShipEnv(65536, GC, EnvConfig, bank=bank, seed=SEED, honour_lidar_config=True, device=dev, validate_actions=False)
        env.reset()
        K = 100
        ms = _single_rank_timed(torch, dev, env, K, 10)
        rate = 65536 * K / (ms * 1e-3)
        res["hard_map_65536"] = {"env_steps_per_s": rate, "launch_ms": ms, "K": K,
                                 "roofline_frac": rate * b_alg(K) / 1e9 / peak}
        env.close()
        del env
    # configs[3]: 1,048,576 envs sharded over the ranks, 128-step rollouts, one stats all-reduce per rollout
    total = 1048576
    off, cnt = sdist.shard(total, rank, world)
    bank = ScenarioBank.generate(N_SCENARIOS, (600, 600), seed=SEED)
    env = BatchedShipEnv(cnt, bank=bank, seed=SEED, device=dev, env_id_offset=off, validate_actions=False)
    env.reset()
    K = 128 if world > 1 else 32       # one GPU: 1M envs x 128 steps of obs would be 17 GB; keep it modest
    ms = timed(env, K, 5, True)
    rate = total * K / (ms * 1e-3)
    res["sharded_1048576"] = {"env_steps_per_s": rate, "launch_ms": ms, "K": K, "envs_per_gpu": cnt,
                              "roofline_frac_per_gpu": rate / world * b_alg(K) / 1e9 / peak}
    # K=1 gym-style stepping of the same batch (one launch per env-step)
    ms1 = timed(env, 1, 50, False)
    rate1 = total / (ms1 * 1e-3)
    res["sharded_1048576_K1"] = {"env_steps_per_s": rate1, "launch_ms": ms1, "K": 1,
                                 "roofline_frac_per_gpu": rate1 / world * b_alg(1) / 1e9 / peak}
    env.close()
    # configs[4]: MLP policy + 16,384 envs on the same GPU, 128-step rollouts replayed from ONE CUDA graph: no host
    # round-trip per step (rank 0 only; nothing here is the headline)
    if rank == 0:
        from ship_sim_gym_b200.rollout import MlpPolicy, RolloutCollector
        n, T = 16384, 128
        env = BatchedShipEnv(n, bank=bank, seed=SEED, device=dev, validate_actions=False)
        torch.manual_seed(SEED)
        col = RolloutCollector(env, MlpPolicy().to(dev), T=T, use_graph=True)
        col.collect()
        col.collect()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            col.collect()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res["policy_rollout_16384"] = {"env_steps_per_s": n * T / (ms * 1e-3), "rollout_ms": ms, "T": T,
                                       "policy": "MlpPolicy 32-64-64 tanh (pi, vf), Gumbel-max sampling, GAE on device",
                                       "host_syncs_per_rollout": 0, "cuda_graph": True}
        env.close()
    return res


def _single_rank_timed(torch, dev, env, K, reps):
    acts = torch.randint(0, 3, (K, env.num_envs), dtype=torch.int32, device=dev)
    out = env.alloc_rollout(K)
    for _ in range(3):
        env.rollout(acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    main()
