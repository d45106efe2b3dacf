"""Sweep of the host-buffer path (shipsim_step_host) at the bench headline shape: envs that go home as complete rows by
DMA x host expansion threads x chunks per rollout, then the self-balancing split.
    python profiles/e2e_sweep.py
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank  # noqa: E402

N, K = 4096, 1000
bank = ScenarioBank.generate(1024, (600, 600), seed=0)
pin = lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()
h_act = pin(K, N, dtype=torch.int32)
h_act.copy_(torch.randint(0, 3, (K, N), dtype=torch.int32))
out = (pin(K, N, 32, dtype=torch.float32).numpy(), pin(K, N, dtype=torch.float32).numpy(), pin(K, N, dtype=torch.uint8).numpy())
print("host cpus", os.cpu_count(), subprocess.run("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA node\\(s\\)'", shell=True, capture_output=True, text=True).stdout)


def timed(env, reps=6):
    for _ in range(2):
        env.step_host(h_act.numpy(), K=K, out=out)
    best, tot = 1e9, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        env.step_host(h_act.numpy(), K=K, out=out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best, tot = min(best, dt), tot + dt
    return tot / reps, best


for threads in (4, 8, 16):
    os.environ["SHIPSIM_HOST_THREADS"] = str(threads)
    env = BatchedShipEnv(N, bank=bank, validate_actions=False)
    env.reset()
    for chunks in (8, 16, 32):
        os.environ["SHIPSIM_HOST_CHUNKS"] = str(chunks)
        for nd in (0, 512, 2048, 3072, 4096):
            os.environ["SHIPSIM_HOST_DMA_ENVS"] = str(nd)
            mean, best = timed(env)
            print("threads=%2d chunks=%2d dma_envs=%4d  mean %.2f ms  best %.2f ms  -> %.3f G env-steps/s (mean)  d2h %.0f MB"
                  % (threads, chunks, nd, mean * 1e3, best * 1e3, N * K / mean / 1e9, env.host_traffic()[1] / 1e6), flush=True)
    del os.environ["SHIPSIM_HOST_DMA_ENVS"], os.environ["SHIPSIM_HOST_CHUNKS"]
    for it in range(12):
        mean, best = timed(env, reps=2)
        print("threads=%2d self-balancing, calls %2d..%2d: mean %.2f ms  -> %.3f G env-steps/s  d2h %.0f MB"
              % (threads, it * 4, it * 4 + 3, mean * 1e3, N * K / mean / 1e9, env.host_traffic()[1] / 1e6), flush=True)
    env.close()
