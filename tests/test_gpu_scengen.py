"""-m gpu: scenario generation on the device (shipsim_generate_scenarios) -- validity of every generated map, the
reference's distributions (against the host generator, which is pinned to reference fixtures), and parity of the
step kernel on a device-generated bank."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from oracle import cbind  # noqa: E402
import parity  # noqa: E402


def _convex_ccw(h):
    a, b, c = np.roll(h, 2, 0), np.roll(h, 1, 0), h
    return (((b - a)[:, 0] * (c - b)[:, 1] - (b - a)[:, 1] * (c - b)[:, 0]) > 0).all()


@pytest.mark.parametrize("W,map_N,wf", [(600, 10, 0.5), (1000, 30, 0.9)])
def test_generated_maps_are_valid_and_distributed_like_the_reference(W, map_N, wf):
    from ship_sim_gym_b200 import BatchedShipEnv, ScenarioBank
    from ship_sim_gym_b200.config import EnvConfig, GameConfig

    class GC(GameConfig):
        BOUNDS = (W, W)
    S = 2048
    env = BatchedShipEnv(64, GC, EnvConfig, n_scenarios=S, seed=7, map_N=map_N, map_width_frac=wf, scenario_source="device")
    dev = env.read_scenarios()
    host = ScenarioBank.generate(S, (W, W), seed=7, map_N=map_N, width_frac=wf)
    bw = wf * W / 2
    assert dev.hull_n.min() >= 3 and dev.hull_n.max() <= 32
    n_checked = 0
    for s in range(S):
        for b in range(2):
            h = dev.hull_xy[s, b, :dev.hull_n[s, b]]
            assert _convex_ccw(h), (s, b)
            assert (h == h.astype(np.float32)).all()                       # the fp32 polygon the kernels use
            x_lo, x_hi = (0, bw) if b == 0 else (W - bw, W)
            assert h[:, 0].min() >= x_lo - 1e-3 and h[:, 0].max() <= x_hi + 1e-3
            assert (W if b else 0) in h[:, 0]                               # the wall corners are on the hull
        g = dev.goals[s]
        assert (np.abs(g[:, 1] - np.arange(1, 6) * W / 6) <= 20 + 1e-9).all()      # y_i = i*H/6 + randint(-20, 20)
        if s % 16 == 0:                                                     # game.py:322-325 against the oracle's fat rays
            for k in range(5):
                ok, lo, hi = cbind.goal_span(dev.hull_xy[s], dev.hull_n[s], float(W), float(g[k, 1]))
                if ok:
                    assert min(lo, hi) - 1e-2 <= g[k, 0] <= max(lo, hi) + 1e-2, (s, k, lo, hi, g[k])
                    n_checked += 1
    assert n_checked > 300
    # distributions: same generator, different random streams -> compare moments (S = 2048 maps)
    def moments(bank):
        left_max = np.array([bank.hull_xy[s, 0, :bank.hull_n[s, 0], 0].max() for s in range(S)])
        right_min = np.array([bank.hull_xy[s, 1, :bank.hull_n[s, 1], 0].min() for s in range(S)])
        return dict(n=bank.hull_n.mean(), left=left_max.mean(), left_sd=left_max.std(), right=right_min.mean(),
                    gx=bank.goals[:, :, 0].mean(), gx_sd=bank.goals[:, :, 0].std(), gy_sd=(bank.goals[:, :, 1] - np.arange(1, 6) * W / 6).std())
    md, mh = moments(dev), moments(host)
    for k in md:
        tol = 0.08 * max(abs(mh[k]), 1.0) if k.endswith("_sd") or k == "n" else 0.02 * W
        assert abs(md[k] - mh[k]) <= tol, (k, md[k], mh[k])
    env.close()


def test_step_parity_on_a_device_generated_bank():
    from ship_sim_gym_b200 import BatchedShipEnv
    n, K, S = 4096, 16, 64
    env = BatchedShipEnv(n, n_scenarios=S, seed=11, scenario_source="device", auto_reset=True)
    bank = env.read_scenarios()
    # goals are stored as fp32 on the device
    bd = dict(hull_xy=bank.hull_xy, hull_n=bank.hull_n, goals=bank.goals.astype(np.float32).astype(np.float64))
    orc = oracle.OracleEnv(n, bd, auto_reset=True, seed=11)
    env.reset()
    orc.reset()
    rng = np.random.RandomState(3)
    st = parity.f32_inputs(*parity.random_states(rng, n, 600, 600, S, bd["goals"], near=(bd["hull_xy"], bd["hull_n"])))
    env.set_state(*st)
    parity.load_oracle_state(orc, *st)
    acts = rng.randint(0, 3, (K, n)).astype(np.int32)
    obs, rew, done = [t.cpu().numpy() for t in env.rollout(torch.tensor(acts, device=env.device))]
    ref = orc.step(acts)
    rep = parity.compare_steps(ref, obs, rew, done, margin_thr=parity.MARGIN_THR, label="device bank")
    assert rep["grazing_frac"] < parity.MAX_EXCLUDED and rep["excluded_frac"] < parity.MAX_EXCLUDED_CUMULATIVE and rep["max_pose_rel_err"] < 2 * parity.REL_TOL
    # regenerating gives a different bank, deterministically
    env.generate_scenarios(S, seed=12)
    b2 = env.read_scenarios()
    env.generate_scenarios(S, seed=11)
    b3 = env.read_scenarios()
    assert not np.array_equal(b2.goals, bank.goals) and np.array_equal(b3.goals, bank.goals) and np.array_equal(b3.hull_xy, bank.hull_xy)
    with pytest.raises(ValueError):
        env.generate_scenarios(S, map_N=31)
    env.close()


def test_fresh_maps_no_env_meets_a_map_twice_and_retired_slices_come_back_new():
    """shipsim_fresh_maps (ShipGame.reset builds a new level per episode, game.py:271-272): resets pick only from the
    newest slice of the device-generated bank, an env never meets the same (slice generation, scenario) twice, slices are
    regenerated once retired (new, valid maps), and nothing is regenerated while an episode that began on it can still run."""
    from ship_sim_gym_b200 import BatchedShipEnv
    from ship_sim_gym_b200.config import EnvConfig, GameConfig

    class EC(EnvConfig):
        MAX_STEPS = 40
    n, S, K = 512, 64, 8
    env = BatchedShipEnv(n, GameConfig, EC, n_scenarios=S, seed=5, scenario_source="device", auto_reset=True, fresh_maps=True)
    q = S // 4
    info = env.fresh_info()
    assert info == {"enabled": True, "period": 0, "pick_base": 0, "pick_count": q, "generation": [0, 0, 0, 0]}
    bank0 = env.read_scenarios()
    env.reset()
    seen = [set() for _ in range(n)]              # per env: (generation of the slice, scenario) of every episode observed
    last_ep = np.full(n, -1)
    periods = set()
    for it in range(90):                          # 720 steps = 18 periods of 40 steps
        env.rollout(None, K=K)
        info = env.fresh_info()
        periods.add(info["period"])
        ints = env.get_state()["ints"]            # rudder, alive, steps, scenario, episode
        scen, ep, steps = ints[:, 3], ints[:, 4], ints[:, 2]
        new = ep != last_ep
        # an episode that began during this launch picked from the slice of this launch's period
        fresh_now = new & (steps < K) & (last_ep >= 0)
        assert ((scen[fresh_now] >= info["pick_base"]) & (scen[fresh_now] < info["pick_base"] + q)).all()
        # nobody sits on the slice that is being regenerated (period >= 3: slice (period + 1) % 4) ...
        if info["period"] >= 3:
            assert ((scen // q) != (info["period"] + 1) % 4).all()
        # ... so the regeneration count of an env's slice identifies the maps it holds: (generation, scenario) = one map
        for e in np.nonzero(new)[0]:
            key = (info["generation"][int(scen[e]) // q], int(scen[e]))
            assert key not in seen[e], (e, key, it)
            seen[e].add(key)
        last_ep = ep.copy()
    assert len(periods) >= 16 and max(len(s) for s in seen) > 20
    info = env.fresh_info()
    assert min(info["generation"]) >= 3
    bank1 = env.read_scenarios()
    assert not np.array_equal(bank1.goals, bank0.goals)
    for sl in range(4):                           # every slice holds new maps; all of them valid
        assert not np.array_equal(bank1.hull_xy[sl * q:(sl + 1) * q], bank0.hull_xy[sl * q:(sl + 1) * q])
    for s in range(S):
        for b in range(2):
            assert 3 <= bank1.hull_n[s, b] <= 12 and _convex_ccw(bank1.hull_xy[s, b, :bank1.hull_n[s, b]])
    # all four slices of one generation are different maps (the seed moves with the period)
    assert len({bank1.goals[sl * q].tobytes() for sl in range(4)}) == 4
    env.fresh_maps(False)
    assert env.fresh_info()["enabled"] is False and env.fresh_info()["pick_count"] == 0
    env.close()
    # needs a device-generated bank of 4 * 2^k scenarios and auto-reset
    from ship_sim_gym_b200 import _abi
    with pytest.raises(_abi.ShipsimError):
        BatchedShipEnv(64, n_scenarios=16, seed=1, fresh_maps=True)                                    # host bank
    with pytest.raises(ValueError):
        BatchedShipEnv(64, n_scenarios=24, seed=1, scenario_source="device", fresh_maps=True)
    with pytest.raises(_abi.ShipsimError):
        BatchedShipEnv(64, n_scenarios=16, seed=1, scenario_source="device", auto_reset=False, fresh_maps=True)


@pytest.mark.parametrize("window", [1, 16])
def test_step_parity_with_fresh_maps(window):
    """Auto-reset trajectories in fresh-maps mode against the oracle given the same pick rule (slice walk)."""
    from ship_sim_gym_b200 import BatchedShipEnv
    n, K, S = 2048, 32, 64
    env = BatchedShipEnv(n, n_scenarios=S, seed=11, scenario_source="device", auto_reset=True, fresh_maps=True, steps_in_flight=window)
    bank = env.read_scenarios()
    bd = dict(hull_xy=bank.hull_xy, hull_n=bank.hull_n, goals=bank.goals.astype(np.float32).astype(np.float64))
    orc = oracle.OracleEnv(n, bd, auto_reset=True, seed=11, pick_base=0, pick_count=S // 4)
    env.reset()
    orc.reset()
    rng = np.random.RandomState(3)
    st = parity.f32_inputs(*parity.random_states(rng, n, 600, 600, S, bd["goals"], near=(bd["hull_xy"], bd["hull_n"])))
    env.set_state(*st)
    parity.load_oracle_state(orc, *st)
    acts = rng.randint(0, 3, (K, n)).astype(np.int32)
    obs, rew, done = [t.cpu().numpy() for t in env.rollout(torch.tensor(acts, device=env.device))]
    assert env.launch_info()["steps_in_flight"] == window
    ref = orc.step(acts)
    rep = parity.compare_steps(ref, obs, rew, done, margin_thr=parity.MARGIN_THR, label="fresh maps")
    assert done.sum() > n // 4                    # most envs were reset at least once: the pick rule was exercised
    assert rep["grazing_frac"] < parity.MAX_EXCLUDED and rep["excluded_frac"] < parity.MAX_EXCLUDED_CUMULATIVE and rep["max_pose_rel_err"] < 2 * parity.REL_TOL
    st2 = env.get_state()["ints"]
    was_reset = st2[:, 4] != np.asarray(st[1])[:, 4]
    assert was_reset.sum() > n // 4 and (st2[was_reset, 3] < S // 4).all()      # whoever was reset sits in slice 0
    env.close()
