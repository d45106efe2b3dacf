"""BASELINE configs[4]: MLP policy + 16,384 envs on one GPU, 128-step rollouts replayed from one CUDA graph.
    python profiles/policy_rollout.py [--envs 16384] [--T 128]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank  # noqa: E402
from ship_sim_gym_b200.rollout import MlpPolicy, RolloutCollector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--T", type=int, default=128)
a = ap.parse_args()
bank = ScenarioBank.generate(1024, (600, 600), seed=0)
env = BatchedShipEnv(a.envs, bank=bank, seed=0, validate_actions=False)
torch.manual_seed(0)
col = RolloutCollector(env, MlpPolicy().cuda(), T=a.T, use_graph=True)
col.collect()
col.collect()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    col.collect()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("envs=%d T=%d rollout_ms=%.3f env_steps_per_s=%.4g  actions %s  mean reward %.4f" % (
    a.envs, a.T, ms, a.envs * a.T / ms * 1e3, torch.bincount(col.actions.flatten(), minlength=3).tolist(), col.rewards.mean().item()))
