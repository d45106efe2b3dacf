"""Soak of shipsim_step_host's streaming pipeline: many full-size calls under varying thread / chunk / DMA-share settings,
every one compared bit for bit with the device-resident rollout of a twin env."""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank
N, K = 4096, 500
bank = ScenarioBank.generate(256, (600, 600), seed=0)
pin = lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()
ho, hr, hd = pin(K, N, 32, dtype=torch.float32), pin(K, N, dtype=torch.float32), pin(K, N, dtype=torch.uint8)
rng = np.random.RandomState(1)
bad = 0
calls = 0
t_start = time.time()
for threads, chunks, dma in itertools.product((2, 5, 16), (None, 3, 64), (None, "0", "1056", "4096")):
    os.environ["SHIPSIM_HOST_THREADS"] = str(threads)
    for k, v in (("SHIPSIM_HOST_CHUNKS", chunks), ("SHIPSIM_HOST_DMA_ENVS", dma)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    a = BatchedShipEnv(N, bank=bank, seed=3, validate_actions=False); a.reset()
    b = BatchedShipEnv(N, bank=bank, seed=3, validate_actions=False); b.reset()
    for it in range(4):
        acts = torch.tensor(rng.randint(0, 3, (K, N)).astype(np.int32))
        o, r, d = a.rollout(acts.cuda())
        ho.fill_(5.0); hr.fill_(5.0); hd.fill_(7)
        b.step_host(acts.numpy(), K=K, out=(ho.numpy(), hr.numpy(), hd.numpy()))
        ok = torch.equal(o.cpu(), ho) and torch.equal(r.cpu(), hr) and torch.equal(d.cpu(), hd)
        calls += 1
        if not ok:
            bad += 1
            print("MISMATCH threads=%s chunks=%s dma=%s call %d" % (threads, chunks, dma, it), flush=True)
    a.close(); b.close()
print("soak: %d calls of %d x %d env-steps, %d mismatches, %.0f s" % (calls, N, K, bad, time.time() - t_start))
sys.exit(1 if bad else 0)
