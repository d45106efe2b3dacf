"""Sweep of the host-buffer path (shipsim_step_host): chunks per rollout x host assembly threads, bench headline shape.
    python profiles/e2e_sweep.py
"""
import os
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
print("host cpus", os.cpu_count())
for threads in (8, 12, 16, 24, 32):
    os.environ["SHIPSIM_HOST_THREADS"] = str(threads)
    env = BatchedShipEnv(N, bank=bank, validate_actions=False)
    env.reset()
    for chunks in (8, 16, 32, 64):
        os.environ["SHIPSIM_HOST_CHUNKS"] = str(chunks)
        for _ in range(2):
            env.step_host(h_act.numpy(), K=K, out=out)
        best, tot = 1e9, 0.0
        for _ in range(6):
            t0 = time.perf_counter()
            env.step_host(h_act.numpy(), K=K, out=out)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best, tot = min(best, dt), tot + dt
        print("threads=%2d chunks=%2d  mean %.2f ms  best %.2f ms  -> %.3f G env-steps/s (mean)" % (threads, chunks, tot / 6 * 1e3, best * 1e3, N * K * 6 / tot / 1e9), flush=True)
    env.close()
