"""Tiny driver for ncu captures: runs `reps` launches of the fused step kernel at a given shape.
    ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -o gpurun_out/prof \
        python profiles/prof_driver.py --envs 4096 --K 1000 --reps 4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank  # noqa: E402
from ship_sim_gym_b200.config import EnvConfig, GameConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--window", type=int, default=0, help="steps_in_flight: 0 auto, 1 serial-in-time kernel, 4/8/16/32 window kernel")
ap.add_argument("--hard", action="store_true")
ap.add_argument("--wf", type=float, default=0.5, help="map width_frac (0.01 = banks hug the walls: almost no env is near a bank)")
ap.add_argument("--presteps", type=int, default=200, help="untimed env-steps first, so envs are spread over their episodes")
a = ap.parse_args()

if a.hard:
    class GC(GameConfig):
        BOUNDS = (1000, 1000)
    bank = ScenarioBank.generate(256, (1000, 1000), seed=0, map_N=30, width_frac=0.9)
    env = BatchedShipEnv(a.envs, GC, EnvConfig, bank=bank, honour_lidar_config=True, lanes_per_env=a.lanes, steps_in_flight=a.window, validate_actions=False)
else:
    bank = ScenarioBank.generate(1024, (600, 600), seed=0, width_frac=a.wf)
    env = BatchedShipEnv(a.envs, bank=bank, lanes_per_env=a.lanes, steps_in_flight=a.window, validate_actions=False)
env.reset()
acts = torch.randint(0, 3, (a.K, a.envs), dtype=torch.int32, device="cuda")
out = env.alloc_rollout(a.K)
pre = 0
while pre < a.presteps:            # reset kernel + these launches come before the captured ones (ncu -s)
    env.rollout(None, K=min(50, a.presteps), out=env.alloc_rollout(min(50, a.presteps)))
    pre += 50
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(a.reps):
    env.rollout(acts, out=out)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / a.reps
print("envs=%d K=%d launch_ms=%.4f env_steps_per_s=%.4g %s" % (a.envs, a.K, ms, a.envs * a.K / ms * 1e3, env.launch_info()))
