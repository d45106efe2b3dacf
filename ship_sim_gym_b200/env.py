"""BatchedShipEnv -- N ShipEnvs resident on one B200, stepped by the fused CUDA kernel through the C ABI.

Mirrors the reference interface for the path (ship_gym/ship_env.py:16-184): `reset()`, `step(actions)`,
`seed()`, `render()`, `action_space` / `observation_space` / `reward_range` / `metadata`, `n_states`,
`states_history`, the `GameConfig` / `EnvConfig` knobs, the same exceptions for the same mistakes
(AssertionError for an action outside Discrete(3), ValueError for HISTORY_SIZE < 1).  PyTorch is used only to
own device memory and streams; all arithmetic happens in libshipsim.so.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _abi
from .config import BASE_DT, SPACE_DAMPING, snapshot
from .scenario import ScenarioBank

STEP_PENALTY = -0.01          # ship_env.py:13
DEFAULT_STATE_VAL = -1        # ship_env.py:12


class Discrete(object):
    """Stand-in for gym.spaces.Discrete (gym is not a dependency): ship_env.py:19."""

    def __init__(self, n):
        self.n = n
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
            return 0 <= int(x) < self.n
        if isinstance(x, np.ndarray) and x.shape == () and x.dtype.kind in "iu":
            return 0 <= int(x) < self.n
        return False

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(object):
    """Stand-in for gym.spaces.Box as the reference declares it (ship_env.py:48).  Note the reference declares
    uint8 in [0, max(bounds)] but returns float64 values including -1; we keep the
    declaration and return float32."""

    def __init__(self, low, high, shape, dtype):
        self.low = np.full(shape, low, dtype=np.float32)
        self.high = np.full(shape, high, dtype=np.float32)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)


def _dtype_code(t):
    if t.dtype == torch.int32:
        return _abi.ACTION_I32
    if t.dtype == torch.int64:
        return _abi.ACTION_I64
    if t.dtype == torch.uint8:
        return _abi.ACTION_U8
    raise TypeError("actions must be int32, int64 or uint8, got %s" % t.dtype)


class BatchedShipEnv(object):
    """`num_envs` independent ShipEnvs on one GPU.

    step(actions) -> (obs [N,16*H] f32, reward [N] f32, done [N] bool, info {})      one env-step per call
    rollout(actions [K,N] | None, K) -> (obs [K,N,16*H], reward [K,N], done [K,N])   K steps in ONE launch
    With auto_reset (default, the SubprocVecEnv worker semantics behind train/stable_baselines/ppo.py:123) a
    done env is reset inside the kernel and the obs returned for that step is the reset obs.
    """

    metadata = {"render.modes": ["human", "rgb_array"]}       # ship_env.py:18
    reward_range = (-1, 1)                                    # ship_env.py:20

    def __init__(self, num_envs, game_config=None, env_config=None, device=None, seed=0, n_scenarios=1024,
                 bank=None, map_N=10, map_width_frac=0.5, auto_reset=True, honour_lidar_config=False,
                 env_id_offset=0, lanes_per_env=0, validate_actions=True, scenario_source="host", steps_in_flight=0,
                 host_threads=0, fresh_maps=False):
        self.knobs = snapshot(game_config, env_config, honour_lidar_config)
        if self.knobs["lidar"]["N_BEAMS"] != _abi.N_BEAMS:
            raise NotImplementedError("N_BEAMS must be 10")
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda") if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _abi.ShipsimError("BatchedShipEnv needs a CUDA device: there is no CPU fallback")
        if self.device.index is None:          # 'cuda' = the CURRENT device, not device 0: handle and buffers must agree
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.L = _abi.load()
        self._needs_reset = True
        self.action_space = Discrete(3)                        # ship_env.py:19
        self.history = self.knobs["history"]
        self.n_states = 2 + 1 + 1 + 2 + _abi.N_BEAMS           # ship_env.py:43
        self.states_history = self.n_states * self.history     # ship_env.py:44
        self.observation_space = Box(0, max(self.knobs["W"], self.knobs["H"]), (self.states_history,), np.uint8)
        self.bounds = (self.knobs["W"], self.knobs["H"])
        self.auto_reset = bool(auto_reset)
        self.validate_actions = bool(validate_actions)
        self.seed_value = int(seed)
        self._kernel_hist = min(self.history, 2)               # longer histories are assembled here from frames

        dt = BASE_DT * self.knobs["speed"]                     # game.py:194
        cfg = _abi.default_config()
        cfg.num_envs = self.num_envs
        cfg.env_id_offset = int(env_id_offset)
        cfg.seed = self.seed_value & 0xFFFFFFFFFFFFFFFF
        cfg.bounds_w, cfg.bounds_h = self.knobs["W"], self.knobs["H"]
        cfg.dt = dt
        cfg.damping = math.pow(SPACE_DAMPING, dt)              # cpSpaceStep: pow(space.damping, dt), in double
        cfg.max_steps = self.knobs["max_steps"]
        cfg.history = self._kernel_hist if self.history <= 2 else 1
        cfg.auto_reset = int(self.auto_reset)
        cfg.lidar_spread_deg = self.knobs["lidar"]["ANGULAR_SPREAD"]
        cfg.lidar_distance = self.knobs["lidar"]["DISTANCE"]
        cfg.lanes_per_env = int(lanes_per_env)
        cfg.steps_in_flight = int(steps_in_flight)     # 0 auto, 1 serial-in-time kernel, 4/8/16/32 time-parallel window
        if not host_threads:
            # step_host's row-assembly pool: the box's cores are shared by the ranks torchrun started on it
            import os
            ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            host_threads = max(1, min(16, (os.cpu_count() or 1) // ranks_here))
        cfg.host_threads = int(host_threads)
        self.cfg = cfg
        self.reset_epoch = 0
        self.last_reset_obs = None
        self.params_epoch = 0                  # bumped whenever kernel parameters change (captured CUDA graphs go stale)
        self._h = C.c_void_p()
        _abi.check(self.L.shipsim_create(C.byref(cfg), self.device.index, C.byref(self._h)))
        self._obs_dim_kernel = _abi.FRAME * cfg.history

        self._n_generated = 0
        if scenario_source not in ("host", "device"):
            raise ValueError("scenario_source must be 'host' or 'device'")
        if bank is None and scenario_source == "device":
            # maps generated by the GPU itself (shipsim_generate_scenarios): no host loop over resets
            self.generate_scenarios(n_scenarios, seed=self.seed_value, map_N=map_N, map_width_frac=map_width_frac)
        else:
            if bank is None:
                bank = ScenarioBank.generate(n_scenarios, self.bounds, seed=self.seed_value, map_N=map_N,
                                             width_frac=map_width_frac)
            self.load_scenarios(bank)

        with torch.cuda.device(self.device):
            nbytes = self.L.shipsim_state_bytes(self._h)
            self.state = torch.zeros(nbytes // 4, dtype=torch.float32, device=self.device).view(_abi.STATE_PLANES, self.num_envs, 4)
            self._stats_slots = torch.zeros(self.L.shipsim_stats_bytes(self._h) // 8, dtype=torch.float64, device=self.device)
            self._stats_out = torch.zeros(_abi.STATS_LEN, dtype=torch.float64, device=self.device)
            _abi.check(self.L.shipsim_bind_state(self._h, self.state.data_ptr(), self._stats_slots.data_ptr(), self._stream()))
        self._hist = None
        self.total_steps = 0
        self._state_bound = True
        self.graph_safe = True                 # False: kernel parameters change between launches (no CUDA-graph replay)
        if fresh_maps:
            self.fresh_maps(True)

    # ------------------------------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.L.shipsim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_scenarios(self, bank):
        """Upload a ScenarioBank (curriculum changes call this between rollouts).  On a live batch every env is then
        RESET: its stored scenario id and goals belong to the old bank (ShipGame.reset builds level and goals together,
        game.py:271-272); `last_reset_obs` holds the reset observations."""
        if tuple(bank.bounds) != tuple(self.bounds):
            raise ValueError("scenario bank was generated for bounds %s, env has %s" % (bank.bounds, self.bounds))
        self.bank = bank
        self._n_scen = len(bank)
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_load_scenarios(self._h, bank.hull_xy.ctypes.data, bank.hull_n.ctypes.data,
                                                     bank.goals.ctypes.data, len(bank), bank.maxv))
        self.params_epoch = getattr(self, "params_epoch", 0) + 1
        self.graph_safe = True                 # (a host bank leaves fresh-maps mode: shipsim_load_scenarios)
        if getattr(self, "_state_bound", False) and not self._needs_reset:
            self.reset()

    def _gen_count(self):
        return self._n_generated

    def generate_scenarios(self, n_scenarios, seed=0, map_N=10, map_width_frac=0.5):
        """Replace the scenario bank by `n_scenarios` maps generated on the device (river banks, hulls, goal paths;
        SURVEY.md §8 f2).  Envs keep the goals of the scenario they started with until their next reset, so call it
        where all envs are reset (e.g. followed by `reset()`)."""
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_generate_scenarios(self._h, int(n_scenarios), int(seed) & 0xFFFFFFFFFFFFFFFF, int(map_N),
                                                         float(map_width_frac), self._stream()))
        self.bank = None
        self._n_generated = int(n_scenarios)
        self._n_scen = int(n_scenarios)
        self.params_epoch = getattr(self, "params_epoch", 0) + 1
        self._needs_reset = True
        self.graph_safe = True                 # (a new bank leaves fresh-maps mode: shipsim_generate_scenarios)

    def fresh_maps(self, enable=True):
        """A new map for every episode, as ShipGame.reset builds one (game.py:271-272): the device-generated bank
        (`scenario_source="device"` or `generate_scenarios`; 4 * 2^k scenarios) is regenerated slice by slice on a side
        stream behind the envs, and resets only pick from the newest slice (shipsim_fresh_maps).  Needs auto_reset."""
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_fresh_maps(self._h, int(bool(enable))))
        self.graph_safe = not enable
        self.params_epoch += 1

    def fresh_info(self):
        """dict(enabled, period, pick_base, pick_count, generation[4]) of the fresh-maps rotation."""
        a = (C.c_int32 * 8)()
        _abi.check(self.L.shipsim_fresh_info(self._h, a))
        return {"enabled": bool(a[0]), "period": a[1], "pick_base": a[2], "pick_count": a[3], "generation": list(a[4:8])}

    def read_scenarios(self):
        """The device-generated bank as a host ScenarioBank (validation / inspection)."""
        S = self._gen_count()
        hull_xy = np.zeros((S, 2, _abi.MAX_HULL, 2))
        hull_n = np.zeros((S, 2), dtype=np.int32)
        goals = np.zeros((S, 5, 2))
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_read_scenarios(self._h, hull_xy.ctypes.data, hull_n.ctypes.data, goals.ctypes.data))
        return ScenarioBank(hull_xy, hull_n, goals, self.bounds)

    def set_max_steps(self, max_steps):
        """Change EnvConfig.MAX_STEPS (config.py:16) of the live batch -- a curriculum knob."""
        _abi.check(self.L.shipsim_set_max_steps(self._h, int(max_steps)))
        self.knobs["max_steps"] = int(max_steps)
        self.params_epoch += 1

    def seed(self, seed=None):
        """ship_env.py:52-60 seeds numpy only; here it also re-keys the action-space sampler."""
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 31))
        np.random.seed(seed % (2 ** 32))
        self.action_space.seed(seed)
        return [seed]

    def render(self, mode="human", close=False, env_index=0, size=None):
        """ship_env.py:158-168 only prints; `rgb_array` (listed in metadata, ship_env.py:18, never implemented there)
        draws ONE env on the device the way ShipGame.render does (game.py:197-229) and returns a uint8 CUDA tensor
        [height, width, 3] (the gym convention); `size` = (width, height), default = BOUNDS."""
        if mode != "rgb_array":
            return None
        if self._needs_reset:
            raise _abi.ShipsimError("call reset() before render()")
        w, h = (int(self.knobs["W"]), int(self.knobs["H"])) if size is None else (int(size[0]), int(size[1]))
        img = torch.empty(h, w, 3, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_render(self._h, int(env_index), w, h, img.data_ptr(), self._stream()))
        return img

    def get_screen(self, env_index=0):
        """ShipGame.get_screen (game.py:133-138): pygame.surfarray.array3d layout, [width, height, 3]."""
        return self.render("rgb_array", env_index=env_index).permute(1, 0, 2).contiguous()

    # ------------------------------------------------------------------------------------------ reset / step
    def reset(self, mask=None, scenario=None):
        """Reset all envs (or those where `mask` is true).  Returns the observation of ALL envs is not
        available for a partial reset -- rows of envs that were not reset are left as zeros."""
        N = self.num_envs
        obs = torch.zeros(N, self._obs_dim_kernel, dtype=torch.float32, device=self.device)
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        sc = None if scenario is None else torch.as_tensor(scenario, device=self.device).to(torch.int32).contiguous()
        if sc is not None:
            if sc.numel() != N:
                raise ValueError("scenario must hold one id per env")
            lo, hi = int(sc.min()), int(sc.max())
            if lo < 0 or hi >= self._n_scen:
                raise ValueError("scenario ids must be in [0, %d), got [%d, %d]" % (self._n_scen, lo, hi))
        first = 1 if (mask is None and self._needs_reset) else 0
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_reset(self._h, None if m is None else m.data_ptr(), None if sc is None else sc.data_ptr(),
                                            first, obs.data_ptr(), self._stream()))
        self._needs_reset = False
        self.reset_epoch += 1
        self.last_reset_obs = obs if mask is None else None
        if self.history > 2:
            frame = obs
            if self._hist is None or mask is None:
                self._hist = torch.full((N, self.states_history), -1.0, dtype=torch.float32, device=self.device)
                self._hist[:, -_abi.FRAME:] = frame
            else:
                mb = m.bool()
                fresh = torch.full_like(self._hist, -1.0)
                fresh[:, -_abi.FRAME:] = frame
                self._hist = torch.where(mb[:, None], fresh, self._hist)
            return self._hist.clone()
        return obs

    def _check_actions(self, actions):
        if self.validate_actions:
            bad = ((actions < 0) | (actions > 2)).any()
            assert not bool(bad), "%r invalid" % (actions,)       # ship_env.py:143

    def step(self, actions):
        """One env-step for every env.  `actions`: int tensor [N] on the env's device (or anything
        torch.as_tensor accepts; host data is copied)."""
        a = torch.as_tensor(actions)
        if a.device != self.device:
            a = a.to(self.device)
        if a.dtype not in (torch.int32, torch.int64, torch.uint8):
            a = a.to(torch.int64)
        a = a.contiguous().view(-1)
        assert a.numel() == self.num_envs, "expected %d actions" % self.num_envs
        self._check_actions(a)
        obs, rew, done = self._launch(a, 1)
        obs, rew, done = obs[0], rew[0], done[0].bool()
        if self.history > 2:
            frame = obs
            shifted = torch.cat([self._hist[:, _abi.FRAME:], frame], dim=1)
            if self.auto_reset:
                fresh = torch.full_like(shifted, -1.0)
                fresh[:, -_abi.FRAME:] = frame
                shifted = torch.where(done[:, None], fresh, shifted)
            self._hist = shifted
            obs = shifted.clone()
        return obs, rew, done, {}

    def rollout(self, actions=None, K=None, out=None):
        """K consecutive env-steps in one kernel launch.  actions: int tensor [K,N] on device, or None for the
        in-kernel random agent (train/random.py:18).  `out` = (obs, reward, done) preallocated tensors."""
        if self.history > 2 and out is not None:
            raise NotImplementedError("rollout(out=...) supports HISTORY_SIZE <= 2 (longer rows are assembled from frames here)")
        if actions is not None:
            a = actions
            if a.device != self.device:
                a = a.to(self.device)
            a = a.contiguous()
            assert a.dim() == 2 and a.shape[1] == self.num_envs
            K = a.shape[0]
            self._check_actions(a)
        else:
            assert K is not None
            a = None
        if self.history > 2:
            frames, rew, done = self._launch(a, K, None)          # the kernel runs in its one-frame mode
            return self._long_history_rows(frames, done), rew, done
        return self._launch(a, K, out)

    def _long_history_rows(self, frames, done):
        """HISTORY_SIZE > 2: rows [frame t-H+1 | ... | frame t] put together from the K frames of a
        rollout and the running history, with -1 for everything older than an env's latest reset (under auto-reset the
        frame of a done step IS the reset frame: ship_env.py:180-184)."""
        K, N, H, F = frames.shape[0], self.num_envs, self.history, _abi.FRAME
        prev = self._hist.view(N, H, F).permute(1, 0, 2)          # [H][N][F], oldest first
        allf = torch.cat([prev, frames], dim=0)                   # frame of step k sits at index H + k
        rows = allf.unfold(0, H, 1)[1:].permute(0, 1, 3, 2)       # [K][N][H][F]: rows[k] = frames k-H+1 .. k
        if self.auto_reset:
            k_idx = torch.arange(K, device=self.device)[:, None]
            last_reset = torch.where(done.bool(), k_idx, torch.full_like(k_idx, -(1 << 30))).cummax(dim=0).values      # [K][N]
            slot_time = k_idx[:, :, None] - torch.arange(H - 1, -1, -1, device=self.device)[None, None, :]             # [K][1][H]
            rows = torch.where((slot_time < last_reset[:, :, None])[..., None], torch.full_like(rows, -1.0), rows)
        rows = rows.reshape(K, N, H * F).contiguous()
        self._hist = rows[-1].clone()
        return rows

    def alloc_rollout(self, K):
        N = self.num_envs
        return (torch.empty(K, N, self._obs_dim_kernel, dtype=torch.float32, device=self.device),
                torch.empty(K, N, dtype=torch.float32, device=self.device),
                torch.empty(K, N, dtype=torch.uint8, device=self.device))

    def _launch(self, a, K, out=None):
        if self._needs_reset:
            raise _abi.ShipsimError("call reset() before step()")
        obs, rew, done = out if out is not None else self.alloc_rollout(K)
        code = _abi.ACTION_RANDOM if a is None else _dtype_code(a)
        # (no torch.cuda.device() context here: shipsim_step selects the handle's device itself, and on the gym-style K = 1
        # path the two extra cudaSetDevice calls were a fifth of the per-launch host time)
        rc = self.L.shipsim_step(self._h, None if a is None else a.data_ptr(), code, K, obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                 torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _abi.check(rc)
        self.total_steps += K * self.num_envs
        return obs, rew, done

    def step_host(self, actions, K=1, out=None):
        """The CPU-caller path: numpy/pinned int32 actions [K,N] in, numpy obs/reward/done out, with the
        host<->device copies inside (shipsim_step_host)."""
        if self.history > 2:          # (rows assembled on the device, then copied: this path is not tuned for long histories)
            o, r, d = self.rollout(torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32).reshape(K, self.num_envs)))
            res = (o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy())
            if out is not None:
                for dst, src in zip(out, res):
                    if dst is not None:
                        dst[...] = src
                return out
            return res
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(K, self.num_envs)
        if self.validate_actions:
            assert ((a >= 0) & (a <= 2)).all(), "%r invalid" % (actions,)
        if out is None:
            out = (np.empty((K, self.num_envs, self._obs_dim_kernel), dtype=np.float32),
                   np.empty((K, self.num_envs), dtype=np.float32), np.empty((K, self.num_envs), dtype=np.uint8))
        obs, rew, done = out           # any of them may be None: that output is then not copied back
        ptr = lambda x: None if x is None else x.ctypes.data    # noqa: E731
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_step_host(self._h, a.ctypes.data, K, ptr(obs), ptr(rew), ptr(done), self._stream()))
        self.total_steps += K * self.num_envs
        return obs, rew, done

    def step_np(self, actions):
        """One env-step for a caller on the CPU (the gym facade, the vector-env adapters): numpy int actions [N] in,
        (obs [N, 16 * history] float32, reward [N] float32, done [N] bool) out -- ONE call into shipsim_step_host with
        page-locked buffers kept on the env; no torch ops, one stream wait.  The returned arrays are the env's buffers:
        valid until the next call (copy them to keep them)."""
        if self.history > 2:
            obs, rew, done, _ = self.step(torch.as_tensor(np.asarray(actions)))
            return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        if getattr(self, "_np_bufs", None) is None:
            pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype).pin_memory()      # noqa: E731
            self._np_keep = (pin(1, self.num_envs, dtype=torch.int32), pin(1, self.num_envs, self._obs_dim_kernel, dtype=torch.float32),
                             pin(1, self.num_envs, dtype=torch.float32), pin(1, self.num_envs, dtype=torch.uint8))
            self._np_bufs = tuple(t.numpy() for t in self._np_keep)
        a, obs, rew, done = self._np_bufs
        a[0, :] = actions
        if self.validate_actions:
            assert ((a >= 0) & (a <= 2)).all(), "%r invalid" % (actions,)
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_step_host(self._h, a.ctypes.data, 1, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, self._stream()))
        self.total_steps += self.num_envs
        return obs[0], rew[0], done[0].view(np.bool_)

    def host_threads(self):
        """Host threads step_host() uses to rebuild observation rows (0 before its first call)."""
        n = C.c_int32()
        _abi.check(self.L.shipsim_host_threads(self._h, C.byref(n)))
        return n.value

    def host_traffic(self):
        """(host->device, device->host) bytes the last step_host() moved over PCIe."""
        a, b = C.c_int64(), C.c_int64()
        _abi.check(self.L.shipsim_host_traffic(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ------------------------------------------------------------------------------------------ stats / state
    def stats_tensor(self, clear=False, out=None):
        """float64[16] device tensor (see shipsim_stat); ready to be all-reduced.  `out`: write into this tensor
        (e.g. StatsReducer.next_buffer()) instead of the env's own."""
        dst = self._stats_out if out is None else out
        assert dst.dtype == torch.float64 and dst.numel() >= _abi.STATS_LEN and dst.is_contiguous()
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_stats_read(self._h, dst.data_ptr(), int(clear), self._stream()))
        return dst

    def stats(self, clear=False):
        v = self.stats_tensor(clear).cpu().tolist()
        d = dict(zip(_abi.STAT_NAMES, v))
        d["steps"] = float(self.total_steps)
        return d

    def get_state(self):
        """dict of numpy arrays: pose [N,6] (x y angle vx vy w), ints [N,5] (rudder, alive_mask, step_count,
        scenario, episode), lidar [N,10], goals [N,5,2], ep_return [N]."""
        N = self.num_envs
        pose = np.zeros((N, 6), dtype=np.float32)
        ints = np.zeros((N, 5), dtype=np.int32)
        lidar = np.zeros((N, 10), dtype=np.float32)
        goals = np.zeros((N, 5, 2), dtype=np.float32)
        ret = np.zeros(N, dtype=np.float32)
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_get_state(self._h, pose.ctypes.data, ints.ctypes.data, lidar.ctypes.data,
                                                goals.ctypes.data, ret.ctypes.data))
        return dict(pose=pose, ints=ints, lidar=lidar, goals=goals, ep_return=ret)

    def set_state(self, pose, ints, lidar, goals, ep_return):
        N = self.num_envs
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(N, 6)
        ints = np.ascontiguousarray(ints, dtype=np.int32).reshape(N, 5)
        lidar = np.ascontiguousarray(lidar, dtype=np.float32).reshape(N, 10)
        goals = np.ascontiguousarray(goals, dtype=np.float32).reshape(N, 10)
        ret = np.ascontiguousarray(ep_return, dtype=np.float32).reshape(N)
        with torch.cuda.device(self.device):
            _abi.check(self.L.shipsim_set_state(self._h, pose.ctypes.data, ints.ctypes.data, lidar.ctypes.data,
                                                goals.ctypes.data, ret.ctypes.data))
        self._needs_reset = False

    def launch_info(self):
        n = C.c_int64()
        lanes, thr, ctas = C.c_int32(), C.c_int32(), C.c_int32()
        _abi.check(self.L.shipsim_launch_count(self._h, C.byref(n)))
        _abi.check(self.L.shipsim_launch_shape(self._h, C.byref(lanes), C.byref(thr), C.byref(ctas)))
        win = C.c_int32()
        _abi.check(self.L.shipsim_launch_window(self._h, C.byref(win)))
        return dict(launches=n.value, lanes_per_env=lanes.value, threads_per_cta=thr.value, ctas=ctas.value,
                    steps_in_flight=win.value)


class ShipEnv(object):
    """Single-env facade with the reference constructor and numpy in/out: `ShipEnv(game_config, env_config)`
    (ship_env.py:23) -- what train/random.py:9-26 and the `make_env()` thunks construct.  It is a batch of one;
    like the reference, `step` does NOT auto-reset."""

    metadata = BatchedShipEnv.metadata
    reward_range = BatchedShipEnv.reward_range
    action_space = Discrete(3)

    def __init__(self, game_config=None, env_config=None, **kw):
        kw.setdefault("n_scenarios", 64)
        kw["auto_reset"] = False
        self.batch = BatchedShipEnv(1, game_config, env_config, **kw)
        self.env_config = env_config
        self.n_states = self.batch.n_states
        self.states_history = self.batch.states_history
        self.observation_space = self.batch.observation_space
        self.last_action = None
        self.reward = 0
        self.cumulative_reward = 0
        self.step_count = 0
        self.episodes_count = -1                               # ship_env.py:33

    def seed(self, seed=None):
        return self.batch.seed(seed)

    def reset(self):
        obs = self.batch.reset(mask=None if self.episodes_count < 0 else torch.ones(1, dtype=torch.uint8))
        self.last_action = None
        self.reward = 0
        self.cumulative_reward = 0
        self.step_count = 0
        self.episodes_count += 1
        return obs[0].cpu().numpy()

    def step(self, action):
        assert self.action_space.contains(action), "%r (%s) invalid" % (action, type(action))   # ship_env.py:143
        obs, rew, done = self.batch.step_np(int(action))
        self.last_action = action
        self.reward = float(rew[0])
        self.cumulative_reward += self.reward
        self.step_count += 1
        return obs[0].copy(), self.reward, bool(done[0]), {}

    def render(self, mode="human", close=False):
        if mode == "rgb_array":
            return self.batch.render("rgb_array").cpu().numpy()
        import sys
        if self.last_action is not None:
            sys.stdout.write("action=%s, cumm_reward=%s" % (self.last_action, self.cumulative_reward))

    def close(self):
        self.batch.close()
