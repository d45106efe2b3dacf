import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank
from ship_sim_gym_b200.rollout import MlpPolicy, RolloutCollector
N, T = 16384, 128
bank = ScenarioBank.generate(1024, (600, 600), seed=0)
env = BatchedShipEnv(N, bank=bank, seed=0, validate_actions=False)
torch.manual_seed(0)
col = RolloutCollector(env, MlpPolicy().cuda(), T=T, use_graph=False)
col.collect(); torch.cuda.synchronize()
def timed(fn, reps=3):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    print("policy kernel x%d: %.3f ms" % (T, timed(lambda: [col._forward_kernel(t) for t in range(T)])))
    print("env step x%d:      %.3f ms" % (T, timed(lambda: [env.rollout(col.actions[t:t + 1], out=(col.obs[t + 1:t + 2], col.rewards[t:t + 1], col.dones[t:t + 1])) for t in range(T)])))
    print("noise:             %.3f ms" % timed(lambda: col._noise.uniform_().clamp_(1e-10, 1.0).log_().neg_().log_().neg_()))
    print("gae:               %.3f ms" % timed(col._gae))
    A = 3
    print("values+logp:       %.3f ms" % timed(lambda: (col.values.copy_(col._out[:, :, A]), col.logp.copy_(torch.log_softmax(col._out[:T, :, :A], dim=-1).gather(-1, col.actions[:, :, None]).squeeze(-1)))))
    print("whole collect:     %.3f ms" % timed(col._collect))
