#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python (imported from
/root/reference) over the restated Chipmunk in oracle/shims/ (pymunk -> minimunk; pygame/gym -> stubs).

TEST INFRASTRUCTURE ONLY.  Run here (the build container): /root/reference does not exist on the GPU
box, so only the committed .npz fixtures travel.  Usage:  python oracle/make_golden.py

What the fixtures pin (and what they do not) is stated in oracle/shims/pymunk/__init__.py: all of the
reference's own logic is executed for real; only the Chipmunk arithmetic underneath is a restatement.

Fixture layout (every array float64 unless noted):
  trajectories.npz    one record per episode, keys prefixed "e{idx}_":
      cfg        [W, H, SPEED, HISTORY_SIZE, MAX_STEPS, map_N, map_width_frac]
      raw0/raw1  raw vertex lists handed to pm.Poly (game_map.gen_river_poly output), (n,2)
      hull0/hull1 the convex hulls the shim built from them (CCW), (m,2)
      goals      (5,2) goal centres in creation order
      actions    int64 (T,)
      obs        (T+1, 16*H): row 0 is reset(), row t+1 is step(actions[t])[0]
      reward     (T,) ; done uint8 (T,) ; colliding uint8 (T,) ; goal_reached uint8 (T,)
      state      (T+1, 21): x y angle vx vy w rudder alive_mask step_count cum_reward lidar[10] n_goals
  scenarios.npz       RNG-order pins: for seed s, `random.seed(s); np.random.seed(s)` then one
      ShipGame.reset() -> raw banks + goals, keys "s{seed}_{W}x{H}_raw0" ...
"""
import functools
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if "pymunk" not in sys.modules:      # oracle/live_check.py imports the REAL pymunk first; everyone else gets the shim
    sys.path.insert(0, os.path.join(HERE, "shims"))
if not any(os.path.isdir(os.path.join(p, "ship_gym")) for p in sys.path if p):
    sys.path.insert(0, os.environ.get("SHIPSIM_REF_PATH", "/root/reference"))

from ship_gym import game_map  # noqa: E402  (the real reference)
from ship_gym.config import EnvConfig, GameConfig  # noqa: E402
from ship_gym.ship_env import ShipEnv  # noqa: E402

_real_gen_river_poly = game_map.gen_river_poly


def _snapshot(env, goals0):
    g = env.game
    b = g.player.body
    alive = 0
    for k, go in enumerate(goals0):
        if any(go is x for x in g.goals):
            alive |= 1 << k
    return ([b.position.x, b.position.y, b.angle, b.velocity.x, b.velocity.y, b.angular_velocity,
             g.player.rudder_angle, alive, env.step_count, env.cumulative_reward]
            + [float(v) for v in g.player.lidar.vals] + [len(g.goals)])


def policy_actions(kind, T, rng):
    if kind == "random":
        return rng.randint(0, 3, size=T)
    if kind == "thrusty":      # mostly thrust, some rudder: long runs that meet banks and goals
        return rng.choice([0, 0, 0, 1, 2], size=T)
    if kind == "straight":
        return np.zeros(T, dtype=np.int64)
    if kind == "left":
        return np.array([1, 1] + [0] * (T - 2))
    if kind == "right":
        return np.array([2, 2] + [0] * (T - 2))
    if kind == "left1":
        return np.array([1] + [0] * (T - 1))
    if kind == "right1":
        return np.array([2] + [0] * (T - 1))
    if kind == "wiggle":       # never thrusts: runs into MAX_STEPS
        return np.array([1, 2] * (T // 2 + 1))[:T]
    raise ValueError(kind)


def run_episode(seed, kind, W=600, H=600, speed=10, hist=2, max_steps=1000, T=200, map_N=10, map_wf=0.5):
    GameConfig.BOUNDS = (W, H)
    GameConfig.SPEED = speed
    GameConfig.FPS = 100000
    GameConfig.DEBUG = False
    EnvConfig.HISTORY_SIZE = hist
    EnvConfig.MAX_STEPS = max_steps
    game_map.gen_river_poly = functools.partial(_real_gen_river_poly, N=map_N, width_frac=map_wf)
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.RandomState(1000 + seed)
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        env = ShipEnv(GameConfig, EnvConfig)
        obs0 = env.reset()
    finally:
        sys.stdout.close()
        sys.stdout = stdout
    g = env.game
    goals0 = list(g.goals)
    rec = {
        "cfg": np.array([W, H, speed, hist, max_steps, map_N, map_wf], dtype=np.float64),
        "raw0": np.array(g.level.poly_list[0], dtype=np.float64),
        "raw1": np.array(g.level.poly_list[1], dtype=np.float64),
        "hull0": np.array([tuple(v) for v in g.level.shapes[0].get_vertices()], dtype=np.float64),
        "hull1": np.array([tuple(v) for v in g.level.shapes[1].get_vertices()], dtype=np.float64),
        "goals": np.array([[go.x, go.y] for go in goals0], dtype=np.float64),
    }
    acts = policy_actions(kind, T, rng)
    obs, rew, done, coll, goal, state = [np.asarray(obs0, dtype=np.float64)], [], [], [], [], [_snapshot(env, goals0)]
    used = []
    for a in acts:
        o, r, d, info = env.step(int(a))
        assert info == {}
        used.append(int(a))
        obs.append(np.asarray(o, dtype=np.float64))
        rew.append(float(r))
        done.append(bool(d))
        coll.append(bool(g.colliding))
        goal.append(bool(g.goal_reached))
        state.append(_snapshot(env, goals0))
        if d:
            break
    rec.update(actions=np.array(used, dtype=np.int64), obs=np.array(obs), reward=np.array(rew),
               done=np.array(done, dtype=np.uint8), colliding=np.array(coll, dtype=np.uint8),
               goal_reached=np.array(goal, dtype=np.uint8), state=np.array(state, dtype=np.float64))
    return rec


def scenario_pin(seed, W, H):
    GameConfig.BOUNDS = (W, H)
    GameConfig.SPEED = 10
    game_map.gen_river_poly = _real_gen_river_poly
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        random.seed(12345)
        np.random.seed(12345)
        from ship_gym.game import ShipGame
        game = ShipGame(GameConfig)        # constructor already resets once; pin the NEXT reset
        random.seed(seed)
        np.random.seed(seed)
        game.reset()
    finally:
        sys.stdout.close()
        sys.stdout = stdout
    return {"raw0": np.array(game.level.poly_list[0]), "raw1": np.array(game.level.poly_list[1]),
            "goals": np.array([[go.x, go.y] for go in game.goals])}


def all_episodes():
    """The recorded episodes, in fixture order (also what oracle/live_check.py replays against the real pymunk)."""
    episodes = []
    # default config (config.py:14-24): BOUNDS 600x600, SPEED 10, HISTORY 2
    for seed in range(6):
        episodes.append(run_episode(seed, "random", T=120))
    for seed in range(6, 14):
        episodes.append(run_episode(seed, "thrusty", T=200))
    for seed, kind in ((20, "straight"), (21, "left"), (22, "right"), (23, "left1"), (24, "right1"),
                       (25, "left"), (26, "right"), (27, "left1"), (28, "right1")):
        episodes.append(run_episode(seed, kind, T=200))
    # train/random.py settings: SPEED 1
    for seed in (30, 31):
        episodes.append(run_episode(seed, "thrusty", speed=1, T=150))
    episodes.append(run_episode(32, "left", speed=1, T=300))
    # stable-baselines script settings (train/stable_baselines/ppo.py:65-69): SPEED 30, 1000x1000
    for seed, kind in ((40, "random"), (41, "thrusty"), (42, "left1"), (43, "straight")):
        episodes.append(run_episode(seed, kind, W=1000, H=1000, speed=30, T=100))
    # rllib script: SPEED 40 (train/rllib/ppo.py:12-14), default bounds
    episodes.append(run_episode(45, "thrusty", speed=40, T=100))
    # history sizes 1 and 3, MAX_STEPS cut-off
    episodes.append(run_episode(50, "thrusty", hist=1, T=100))
    episodes.append(run_episode(51, "thrusty", hist=3, T=100))
    episodes.append(run_episode(52, "wiggle", max_steps=25, T=100))
    # builder-defined "max difficulty" map: N=30, width_frac=0.9, 1000x1000
    for seed, kind in ((60, "thrusty"), (61, "left1"), (62, "right"), (63, "random"), (64, "straight")):
        episodes.append(run_episode(seed, kind, W=1000, H=1000, speed=10, T=150, map_N=30, map_wf=0.9))
    return episodes


def main():
    if "--live" in sys.argv:             # diff the committed fixtures against the real pymunk instead of regenerating
        import live_check
        live_check.run()
        return
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    episodes = all_episodes()

    flat = {}
    for i, rec in enumerate(episodes):
        for k, v in rec.items():
            flat["e%d_%s" % (i, k)] = v
    flat["n_episodes"] = np.array(len(episodes))
    np.savez_compressed(os.path.join(out_dir, "trajectories.npz"), **flat)

    scen = {}
    for seed in range(8):
        for (W, H) in ((600, 600), (1000, 1000)):
            r = scenario_pin(seed, W, H)
            for k, v in r.items():
                scen["s%d_%dx%d_%s" % (seed, W, H, k)] = v
    np.savez_compressed(os.path.join(out_dir, "scenarios.npz"), **scen)

    # summary for the log
    n_steps = sum(len(r["actions"]) for r in episodes)
    n_coll = sum(int(r["colliding"].any()) for r in episodes)
    n_goal = sum(int(r["goal_reached"].sum()) for r in episodes)
    n_lidar = sum(int((r["state"][:, 10:20] >= 0).any()) for r in episodes)
    print("episodes=%d steps=%d with_collision=%d goals_reached=%d with_lidar_hit=%d"
          % (len(episodes), n_steps, n_coll, n_goal, n_lidar))
    for i, r in enumerate(episodes):
        print(i, "T=%d" % len(r["actions"]), "done=%d" % r["done"][-1], "coll=%d" % r["colliding"].any(),
              "goals=%d" % r["goal_reached"].sum(), "lidar=%d" % (r["state"][:, 10:20] >= 0).any(),
              "final=(%.1f,%.1f)" % (r["state"][-1, 0], r["state"][-1, 1]))


if __name__ == "__main__":
    main()
