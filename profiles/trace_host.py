import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank
N, K = 4096, 1000
bank = ScenarioBank.generate(1024, (600, 600), seed=0)
pin = lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()
h_act = pin(K, N, dtype=torch.int32); h_act.copy_(torch.randint(0, 3, (K, N), dtype=torch.int32))
out = (pin(K, N, 32, dtype=torch.float32).numpy(), pin(K, N, dtype=torch.float32).numpy(), pin(K, N, dtype=torch.uint8).numpy())
os.environ["SHIPSIM_HOST_DMA_ENVS"] = "0"
env = BatchedShipEnv(N, bank=bank, validate_actions=False); env.reset()
for i in range(4):
    t0 = time.perf_counter(); env.step_host(h_act.numpy(), K=K, out=out); torch.cuda.synchronize(); print("call %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
