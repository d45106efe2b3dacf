"""On-device rollout collection: policy and envs on the same GPU, no host round-trip per step (BASELINE.json
configs[4]).

The policy is the reference's: stable-baselines `MlpPolicy` as used at train/stable_baselines/ppo.py:88 -- two tanh
layers of 64 units each for the policy and for the value function -- here as a plain torch module.  One rollout =
T x (policy forward -> categorical sample -> fused env step); the env kernel reads the int64 action tensor the sampler
produced and writes obs / reward / done straight into the rollout buffers.  Because `shipsim_step` allocates nothing
and never synchronises, the whole rollout is captured once in a CUDA graph and replayed.
"""
import torch
import torch.nn as nn

from .env import BatchedShipEnv


class MlpPolicy(nn.Module):
    """stable-baselines MlpPolicy: pi 32 -> 64 -> 64 -> 3, vf 32 -> 64 -> 64 -> 1, tanh (separate trunks)."""

    def __init__(self, obs_dim=32, n_actions=3, hidden=64, obs_scale=1.0 / 600.0):
        super().__init__()
        self.obs_scale = obs_scale

        def trunk(out):
            return nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, out))
        self.pi = trunk(n_actions)
        self.vf = trunk(1)

    def forward(self, obs):
        x = obs * self.obs_scale
        return self.pi(x), self.vf(x).squeeze(-1)


class RolloutCollector(object):
    """Collects T-step rollouts of a BatchedShipEnv under a policy, entirely on the device."""

    def __init__(self, env, policy, T=128, gamma=0.99, lam=0.95, use_graph=True):
        assert isinstance(env, BatchedShipEnv) and env.history <= 2
        self.env, self.policy, self.T, self.gamma, self.lam = env, policy, int(T), gamma, lam
        N, D, dev = env.num_envs, env.states_history, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.empty(T + 1, N, D, **f32)              # obs[t] is what the policy sees at step t
        self.actions = torch.empty(T, N, dtype=torch.int64, device=dev)
        self.logp = torch.empty(T, N, **f32)
        self.values = torch.empty(T + 1, N, **f32)
        self.rewards = torch.empty(T, N, **f32)
        self.dones = torch.empty(T, N, dtype=torch.uint8, device=dev)
        self.adv = torch.empty(T, N, **f32)
        self.returns = torch.empty(T, N, **f32)
        self.use_graph = bool(use_graph)
        self._graph = None
        env.validate_actions = False                            # the range check would need a host sync per step
        self.obs[0].copy_(env.reset())

    @torch.no_grad()
    def _collect(self):
        env, T = self.env, self.T
        for t in range(T):
            logits, v = self.policy(self.obs[t])
            # categorical sample by Gumbel-max: no host sync, graph-capturable
            u = torch.rand_like(logits).clamp_(1e-10, 1.0)
            a = torch.argmax(logits - torch.log(-torch.log(u)), dim=-1)
            self.actions[t].copy_(a)
            self.logp[t].copy_(torch.log_softmax(logits, dim=-1).gather(-1, a[:, None]).squeeze(-1))
            self.values[t].copy_(v)
            env.rollout(self.actions[t:t + 1], out=(self.obs[t + 1:t + 2], self.rewards[t:t + 1], self.dones[t:t + 1]))
        _, v = self.policy(self.obs[T])
        self.values[T].copy_(v)
        # GAE(lambda), as PPO2 computes it
        last = torch.zeros_like(self.values[0])
        for t in range(T - 1, -1, -1):
            nonterminal = 1.0 - self.dones[t].float()
            delta = self.rewards[t] + self.gamma * self.values[t + 1] * nonterminal - self.values[t]
            last = delta + self.gamma * self.lam * nonterminal * last
            self.adv[t].copy_(last)
        torch.add(self.adv, self.values[:T], out=self.returns)

    def collect(self):
        """One rollout; the buffers (obs, actions, logp, values, rewards, dones, adv, returns) hold the result."""
        if not self.use_graph:
            self._collect()
        else:
            if self._graph is None:
                s = torch.cuda.Stream(device=self.env.device)
                s.wait_stream(torch.cuda.current_stream(self.env.device))
                with torch.cuda.stream(s):
                    self._collect()                             # warm-up outside capture (cuBLAS handles, allocator)
                    self.obs[0].copy_(self.obs[self.T])
                torch.cuda.current_stream(self.env.device).wait_stream(s)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._collect()
                    self.obs[0].copy_(self.obs[self.T])         # next rollout continues where this one ended
                return self                                     # the capture pass does not execute; replay below next call
            self._graph.replay()
            return self
        self.obs[0].copy_(self.obs[self.T])
        return self
