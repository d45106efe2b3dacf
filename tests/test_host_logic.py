"""CPU only: the host-side mirror of the reference interface -- config knobs (ship_gym/config.py:8-24), the
curriculum helper (ship_gym/curriculum.py:23-50) and the gym-space stand-ins (ship_env.py:19,48)."""
import numpy as np
import pytest

from ship_sim_gym_b200 import config, curriculum


def test_config_defaults_are_the_reference_values():
    assert config.GameConfig.SPEED == 10 and config.GameConfig.BOUNDS == (600, 600) and config.GameConfig.FPS == 1000
    assert config.EnvConfig.HISTORY_SIZE == 2 and config.EnvConfig.MAX_STEPS == 1000
    assert (config.LidarConfig.N_BEAMS, config.LidarConfig.DISTANCE, config.LidarConfig.ANGULAR_SPREAD) == (10, 100, 180)
    k = config.snapshot()
    assert (k["W"], k["H"], k["speed"], k["history"], k["max_steps"]) == (600.0, 600.0, 10.0, 2, 1000)
    # the reference builds LiDAR with its constructor defaults and never reads LidarConfig (models.py:29,149-150)
    assert k["lidar"] == dict(N_BEAMS=10, DISTANCE=100, ANGULAR_SPREAD=90)
    assert config.snapshot(honour_lidar_config=True)["lidar"]["ANGULAR_SPREAD"] == 180.0


def test_config_classes_are_used_as_mutable_singletons_but_snapshotted():
    class GC(config.GameConfig):
        pass
    GC.SPEED = 30                      # train/stable_baselines/ppo.py:65-69 idiom
    GC.BOUNDS = (1000, 1000)
    k = config.snapshot(GC, config.EnvConfig)
    GC.SPEED = 1
    assert k["speed"] == 30.0 and (k["W"], k["H"]) == (1000.0, 1000.0)
    assert config.BASE_DT * k["speed"] == pytest.approx(3.0)          # game.py:27,194


def test_history_size_must_be_positive():
    class EC(config.EnvConfig):
        HISTORY_SIZE = 0
    with pytest.raises(ValueError):                                   # ship_env.py:46-47
        config.snapshot(None, EC)


def test_curriculum_semantics_incl_quirks():
    c = curriculum.Curriculum([100, 200, 400], [0.2, 0.5], repeat_condition=1)
    assert int(c) == 100 and float(c) == 100.0
    assert not c.progress(0.2)                 # strict '>' (curriculum.py:44)
    assert not c.progress(0.3)                 # first pass: counter 1, needs repeat_condition + 1 passes
    assert not c.progress(0.1)                 # a failing call does NOT reset the counter (quirk Q27)
    assert c.progress(0.3) and int(c) == 200   # second pass advances
    assert not c.progress(0.9)
    assert c.progress(0.9) and int(c) == 400
    assert not c.progress(5.0) and int(c) == 400          # no lesson left
    lesson = curriculum.Lesson({"reward": 0.5, "steps": 10})
    assert lesson.pass_lesson({"reward": 0.5, "steps": 11}) and not lesson.pass_lesson({"reward": 0.4, "steps": 11})
    assert curriculum.LessonCondition.STEPS.value == 0 and curriculum.LessonCondition.REWARD.value == 1


def test_space_stand_ins():
    pytest.importorskip("torch")
    from ship_sim_gym_b200.env import Box, Discrete
    d = Discrete(3)                                                   # ship_env.py:19
    assert d.contains(0) and d.contains(np.int64(2)) and not d.contains(3) and not d.contains(-1)
    assert not d.contains(1.0) and not d.contains(True)
    d.seed(0)
    assert all(0 <= d.sample() < 3 for _ in range(50))
    b = Box(0, 600, (32,), np.uint8)                                  # ship_env.py:48
    assert b.shape == (32,) and b.high.max() == 600 and b.dtype == np.uint8


class _FakeEnv(object):
    """Duck-typed stand-in for BatchedShipEnv: what CurriculumDriver touches."""

    def __init__(self):
        import torch
        self.t = torch.zeros(16, dtype=torch.float64)
        self.max_steps = None
        self.loaded = []

    def stats_tensor(self, clear=False):
        out = self.t.clone()
        if clear:
            self.t.zero_()
        return out

    def set_max_steps(self, n):
        self.max_steps = n

    def load_scenarios(self, bank):
        self.loaded.append(bank)

    def finish(self, episodes, mean_return):
        self.t[0] += episodes
        self.t[1] += episodes * mean_return


def test_curriculum_driver_schedules_max_steps_and_bank_tiers():
    pytest.importorskip("torch")
    env = _FakeEnv()
    cur = curriculum.Curriculum([100, 300, 1000], [-0.5, 0.5], repeat_condition=0)
    drv = curriculum.CurriculumDriver(env, cur, knob="max_steps", min_episodes=10)
    assert env.max_steps == 100                                   # lesson 0 applied at construction
    env.finish(5, 1.0)
    assert drv.update() == (False, None)                          # too few episodes: statistics keep accumulating
    env.finish(5, 1.0)
    adv, mean = drv.update()
    assert adv and mean == pytest.approx(1.0) and env.max_steps == 300
    assert float(env.t[0]) == 0.0                                 # statistics were consumed
    env.finish(20, 0.2)
    assert drv.update() == (False, pytest.approx(0.2)) and env.max_steps == 300
    env.finish(20, 0.9)
    assert drv.update()[0] and env.max_steps == 1000
    # the all-reduce hook sees the vector of this rank and returns the global one
    env2 = _FakeEnv()
    drv2 = curriculum.CurriculumDriver(env2, curriculum.Curriculum([0, 1], [0.0], repeat_condition=0), knob="bank",
                                       banks=["easy", "hard"], all_reduce=lambda t: t * 2)
    assert env2.loaded == ["easy"]
    env2.finish(4, 0.5)
    adv, mean = drv2.update()
    assert adv and mean == pytest.approx(0.5) and env2.loaded == ["easy", "hard"]
    with pytest.raises(ValueError):
        curriculum.CurriculumDriver(env2, cur, knob="bank")


@pytest.mark.parametrize("align", [0, 4])
def test_history_rows_from_frames_host_code(align):
    """shipsim_assemble_history (the host half of shipsim_step_host, plain C++ with streaming stores): observation rows
    [previous frame | frame] rebuilt from frames, reset rows included, for aligned (AVX-512 / SSE2 paths) and
    unaligned (memcpy path) buffers.  No device involved."""
    from ship_sim_gym_b200 import _abi
    L = _abi.load()
    rng = np.random.RandomState(0)
    N, K = 37, 53
    frames = rng.randn((K + 1) * N * 16).astype(np.float32)
    cut = (rng.rand(K * N) < 0.1).astype(np.uint8)
    want = np.empty((K * N, 32), dtype=np.float32)
    f = frames.reshape(K + 1, N, 16)
    want[:, :16] = f[:-1].reshape(-1, 16)
    want[:, 16:] = f[1:].reshape(-1, 16)
    want[cut != 0, :16] = -1.0
    raw = np.zeros(K * N * 32 + 32, dtype=np.float32)
    base = raw.ctypes.data
    off = ((-base) % 64) // 4 + align // 4            # 64-byte aligned start, optionally knocked off by `align` bytes
    out = raw[off:off + K * N * 32]
    fr_raw = np.zeros(frames.size + 32, dtype=np.float32)
    foff = ((-fr_raw.ctypes.data) % 64) // 4
    fr = fr_raw[foff:foff + frames.size]
    fr[:] = frames
    _abi.check(L.shipsim_assemble_history(out.ctypes.data, fr.ctypes.data, cut.ctypes.data, K * N, N))
    assert np.array_equal(out.reshape(-1, 32), want)
    _abi.check(L.shipsim_assemble_history(out.ctypes.data, fr.ctypes.data, None, K * N, N))     # no cut: plain history
    want[:, :16] = f[:-1].reshape(-1, 16)
    assert np.array_equal(out.reshape(-1, 32), want)
    with pytest.raises(ValueError):
        _abi.check(L.shipsim_assemble_history(None, fr.ctypes.data, None, 1, N))


def _encode_delta(frames, rew, done, step_penalty, rng):
    """numpy statement of compact_frames_kernel's wire format (include/shipsim.h: shipsim_expand_delta); the value
    blocks are laid out in a shuffled order, as the device's atomics may."""
    K, N = rew.shape
    nblk = (N + 31) // 32
    bits = frames.view(np.uint32)
    rec = np.zeros((K, N, 4), dtype=np.uint32)
    off = np.zeros((K, nblk), dtype=np.uint32)
    chunks, pos = [], 0
    order = [(k, b) for k in range(K) for b in range(nblk)]
    rng.shuffle(order)
    for k, b in order:
        off[k, b] = pos
        for e in range(b * 32, min(N, b * 32 + 32)):
            ch = bits[k + 1, e, 4:] != bits[k, e, 4:]
            mask = int(sum(1 << j for j in range(12) if ch[j]))
            r = rew[k, e]
            rcode = 1 if r == 1.0 else (2 if r == -1.0 else 0)
            rud = int(frames[k + 1, e, 2]) // 5 + 2
            rec[k, e] = (bits[k + 1, e, 0], bits[k + 1, e, 1], bits[k + 1, e, 3], rud | rcode << 3 | int(done[k, e]) << 5 | mask << 8)
            chunks.append(frames[k + 1, e, 4:][ch])
            pos += int(ch.sum())
    var = np.concatenate(chunks + [np.zeros(1, np.float32)]).astype(np.float32)
    return rec, off, var


@pytest.mark.parametrize("align,history,cut", [(0, 2, 1), (16, 2, 1), (4, 2, 0), (0, 1, 1)])
def test_compacted_frames_expand_to_the_same_rows(align, history, cut):
    """shipsim_expand_delta (host half of the compacted wire format of shipsim_step_host): 16-byte records + changed
    values expand to exactly the rows shipsim_assemble_history builds from plain frames -- reset rows, rewards and done
    flags included -- for ragged env counts (last block partly filled) and all three store paths.  No device."""
    from ship_sim_gym_b200 import _abi
    L = _abi.load()
    rng = np.random.RandomState(1)
    N, K, pen = 75, 41, np.float32(-0.01)
    frames = np.empty((K + 1, N, 16), dtype=np.float32)
    frames[0] = rng.randn(N, 16)
    frames[0, :, 2] = rng.choice([-10, -5, 0, 5, 10], N)
    for k in range(K):                                   # pose changes every step, the other slots now and then
        frames[k + 1] = frames[k]
        frames[k + 1, :, [0, 1, 3]] = rng.randn(3, N)
        frames[k + 1, :, 2] = rng.choice([-10, -5, 0, 5, 10], N)
        ch = rng.rand(N, 12) < 0.15
        frames[k + 1, :, 4:][ch] = rng.randn(int(ch.sum()))
    frames[3, 5, 4] = -0.0                               # bitwise change, equal as floats
    frames[2, 5, 4] = 0.0
    done = (rng.rand(K, N) < 0.1).astype(np.uint8)
    rew = np.where(done != 0, rng.choice([1.0, -1.0, float(pen)], (K, N)), rng.choice([1.0, float(pen)], (K, N), p=[0.1, 0.9])).astype(np.float32)
    rec, off, var = _encode_delta(frames, rew, done, pen, rng)
    want = np.empty((K, N, 16 * history), dtype=np.float32)
    want[:, :, -16:] = frames[1:]
    if history == 2:
        want[:, :, :16] = frames[:-1]
        if cut:
            want[done != 0, :16] = -1.0

    def aligned(n_floats, knock=0):
        raw = np.zeros(n_floats + 32, dtype=np.float32)
        o = ((-raw.ctypes.data) % 64) // 4 + knock // 4
        return raw[o:o + n_floats]
    out = aligned(want.size, align)
    cur = aligned(N * 16)
    cur[:] = frames[0].ravel()
    r_out, d_out = np.zeros((K, N), np.float32), np.zeros((K, N), np.uint8)
    _abi.check(L.shipsim_expand_delta(out.ctypes.data, r_out.ctypes.data, d_out.ctypes.data, rec.ctypes.data, off.ctypes.data, var.ctypes.data,
                                      cur.ctypes.data, K, N, float(pen), cut, history))
    assert np.array_equal(out.view(np.uint32).reshape(want.shape), want.view(np.uint32))
    assert np.array_equal(r_out, rew) and np.array_equal(d_out, done)
    assert np.array_equal(cur.view(np.uint32).reshape(N, 16), frames[K].view(np.uint32))      # running frames = the last step's
    # outputs are optional; the running frames still advance
    cur[:] = frames[0].ravel()
    _abi.check(L.shipsim_expand_delta(None, None, None, rec.ctypes.data, off.ctypes.data, var.ctypes.data, cur.ctypes.data, K, N, float(pen), cut, history))
    assert np.array_equal(cur.view(np.uint32).reshape(N, 16), frames[K].view(np.uint32))
    with pytest.raises(ValueError):
        _abi.check(L.shipsim_expand_delta(out.ctypes.data, None, None, None, off.ctypes.data, var.ctypes.data, cur.ctypes.data, K, N, float(pen), cut, history))
    with pytest.raises(ValueError):
        _abi.check(L.shipsim_expand_delta(out.ctypes.data, None, None, rec.ctypes.data, off.ctypes.data, var.ctypes.data, cur.ctypes.data, K, N, float(pen), cut, 3))


def test_rollout_collector_fused_policy_and_gae_on_cpu():
    """RolloutCollector runs the two trunks of MlpPolicy as one block-diagonal network (observation scale folded into
    layer 1) and computes GAE from deltas of all steps at once: both agree with the plain formulation (train/
    stable_baselines/ppo.py:88 MlpPolicy; PPO2's GAE recurrence).  Host logic only: no env, no device."""
    torch = pytest.importorskip("torch")
    from ship_sim_gym_b200.rollout import MlpPolicy, RolloutCollector
    torch.manual_seed(0)
    pol = MlpPolicy()
    N, D, T, H, A = 257, 32, 9, 64, 3
    c = RolloutCollector.__new__(RolloutCollector)               # the buffers _refresh_fused / _forward_fused touch, without an env
    c.policy = pol
    c._alloc_fused(N, D, H, A, T, dict(dtype=torch.float32))
    c.obs = (torch.rand(T + 1, N, D) * 600.0)
    with torch.no_grad():
        c._refresh_fused()
        for t in (0, T):
            c._forward_fused(t)
            logits, v = pol(c.obs[t])
            assert torch.allclose(c._out[t, :, :A], logits, atol=2e-6) and torch.allclose(c._out[t, :, A], v, atol=2e-6)
    # GAE: the vectorised deltas + one fused multiply-add per step against the textbook loop
    g, lam = 0.99, 0.95
    rew, val = torch.randn(T, N), torch.randn(T + 1, N)
    done = (torch.rand(T, N) < 0.1).to(torch.uint8)
    want, last = torch.empty(T, N), torch.zeros(N)
    for t in range(T - 1, -1, -1):
        nt = 1.0 - done[t].float()
        last = rew[t] + g * val[t + 1] * nt - val[t] + g * lam * nt * last
        want[t] = last
    c.T, c.gamma, c.lam = T, g, lam
    c.rewards, c.values, c.dones = rew, val, done
    c._nonterm, c._coef, c._delta = torch.empty(T, N), torch.empty(T, N), torch.empty(T, N)
    c.adv, c.returns = torch.empty(T, N), torch.empty(T, N)
    c._gae()
    assert torch.allclose(c.adv, want, atol=1e-5) and torch.allclose(c.returns, want + val[:T], atol=1e-5)


def test_clamped_additions_compose_like_the_window_kernel_assumes():
    """window_kernel (shipsim_window.cu, scan 1a) gets the T rudder values of a window from a log2(T)-round prefix over
    (a, lo, hi) triples instead of a T-step chain, relying on: clamp(clamp(r + a, lo, hi) + a', lo', hi') ==
    clamp(r + a + a', max(lo + a', lo'), min(max(hi + a', lo'), hi')) with clamp(x, lo, hi) = min(max(x, lo), hi)
    (Ship.rotate / clamp_rudder, models.py:136-146: rudder steps of 5 clamped to +-10).  Exhaustive over short
    sequences, random over window-sized ones."""
    import itertools
    rng = np.random.RandomState(0)

    def serial(r0, incs):
        out, r = [], r0
        for a in incs:
            r = min(max(r + a, -10), 10)
            out.append(r)
        return out

    def prefix(r0, incs):                      # Kogge-Stone, exactly as the kernel: lanes t >= off take lane t - off first
        f = [(a, -10, 10) for a in incs]
        off = 1
        while off < len(f):
            g = list(f)
            for t in range(off, len(f)):
                pa, plo, phi = f[t - off]
                fa, flo, fhi = f[t]
                g[t] = (pa + fa, max(plo + fa, flo), min(max(phi + fa, flo), fhi))
            f, off = g, off * 2
        return [min(max(r0 + a, lo), hi) for a, lo, hi in f]

    for T in (1, 2, 3, 4):
        for r0 in (-10, -5, 0, 5, 10):
            for incs in itertools.product((-5, 0, 5), repeat=T):
                assert serial(r0, incs) == prefix(r0, list(incs))
    for T in (4, 8, 16, 32):
        for _ in range(500):
            r0 = int(rng.choice([-10, -5, 0, 5, 10]))
            incs = [int(v) for v in rng.choice([-5, 0, 5], size=T)]
            assert serial(r0, incs) == prefix(r0, incs)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm: the float64 restatement on the host cores) runs without a GPU and prints ONE
    JSON line with the keys the driver reads (metric / unit / value, impl, cpu_baseline, e2e with zero copy bytes)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SHIPSIM_BENCH_REF_BUDGET_S="0.5")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["envs_per_gpu"] == 4096


def test_tile_images_layout():
    """ShipVecEnv.render('rgb_array') tiles the per-env pictures like stable-baselines' tile_images: near-square grid,
    row-major, black padding."""
    pytest.importorskip("torch")
    from ship_sim_gym_b200.adapters import tile_images
    imgs = [np.full((4, 6, 3), i + 1, dtype=np.uint8) for i in range(7)]
    t = tile_images(imgs)
    assert t.shape == (3 * 4, 3 * 6, 3) and t.dtype == np.uint8
    for i in range(7):
        r, c = divmod(i, 3)
        assert (t[r * 4:(r + 1) * 4, c * 6:(c + 1) * 6] == i + 1).all()
    assert (t[8:, 6:] == 0).all()
    assert tile_images(imgs[:1]).shape == (4, 6, 3) and tile_images(imgs[:2]).shape == (8, 6, 3)
