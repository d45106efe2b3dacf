import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_sim_gym_b200 import ShipEnv
from ship_sim_gym_b200.config import GameConfig, EnvConfig
from ship_sim_gym_b200.adapters import ShipVecEnv
env = ShipEnv(GameConfig, EnvConfig)
env.reset()
rng = np.random.RandomState(0)
for n in (200, 3000):
    t0 = time.perf_counter()
    for i in range(n):
        o, r, d, _ = env.step(int(rng.randint(0, 3)))
        if d: env.reset()
    dt = time.perf_counter() - t0
print("ShipEnv facade (gym API, 1 env): %.0f steps/s (%.1f us per step)" % (n / dt, dt / n * 1e6))
for N in (16, 1024):
    venv = ShipVecEnv(num_envs=N)
    venv.reset()
    for n in (50, 1000):
        t0 = time.perf_counter()
        for i in range(n):
            venv.step(rng.randint(0, 3, N))
        dt = time.perf_counter() - t0
    print("ShipVecEnv (SB VecEnv API, %d envs): %.0f env-steps/s (%.1f us per call)" % (N, N * n / dt, dt / n * 1e6))
