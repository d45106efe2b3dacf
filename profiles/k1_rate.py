import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank
bank = ScenarioBank.generate(1024, (600, 600), seed=0)
for N in (4096, 16384):
    env = BatchedShipEnv(N, bank=bank, validate_actions=False); env.reset()
    acts = torch.randint(0, 3, (1, N), dtype=torch.int32, device="cuda")
    out = env.alloc_rollout(1)
    for _ in range(200): env.rollout(acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(2000): env.rollout(acts, out=out)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("N=%d K=1 eager: %.2f us per launch (device), host issue %.2f us" % (N, e0.elapsed_time(e1) / 2000 * 1e3, (t1 - t0) / 2000 * 1e6))
    env.close()
