"""CPU: what pins the float64 oracle BESIDES the golden trajectories (SURVEY.md §8c "what pins the restatement"):

1. the analytic known answers of SURVEY.md Appendix D (integrator, damping, thrust point, moment) -- derived from
   Chipmunk's documented cpBodyUpdatePosition / cpBodyUpdateVelocity / cpMomentForPoly, not measured on pymunk;
2. geometric self-consistency of the restated Chipmunk queries against independent brute-force formulations:
   segment-vs-convex-polygon (cpPolyShapeSegmentQuery) against Cyrus-Beck clipping, polygon-vs-polygon contact against
   exhaustive edge intersection / containment, point-to-polygon distance against the edge-by-edge minimum.
The oracle is test infrastructure; nothing here touches the CUDA path."""
import numpy as np
import pytest

import oracle
from oracle import cbind
from ship_sim_gym_b200 import ScenarioBank


def _env(speed, W=600.0, H=600.0, n=1, seed=0, **kw):
    bank = ScenarioBank.generate(4, (W, H), seed=seed).as_dict()
    env = oracle.OracleEnv(n, bank, W=W, H=H, speed=speed, **kw)
    env.reset(scen=[0] * n)
    return env


# ---------------------------------------------------------------------------------------------- Appendix D
APPENDIX_D = {          # SPEED -> y reported after steps 1.., straight thrust from rest (action 0 every step)
    1: [25.0, 25.2, 25.582489, 26.131488, 26.832419, 27.671979],
    10: [25.0, 45.0, 73.0, 104.2, 136.68, 169.672],
    30: [25.0, 205.0, 396.52, 588.77728],
    40: [25.0, 345.0, 673.192],
}


@pytest.mark.parametrize("speed", sorted(APPENDIX_D))
def test_straight_thrust_known_answers(speed):
    W = 1000.0 if speed >= 30 else 600.0
    env = _env(speed, W, W)
    ys = APPENDIX_D[speed]
    out = env.step(np.zeros((len(ys), 1), dtype=np.int32))
    obs = out["obs"][:, 0, 16:]
    assert np.allclose(obs[:, 1], ys, rtol=0, atol=2e-6)            # y
    assert np.all(obs[:, 0] == W / 2) and np.all(obs[:, 3] == 0.0)  # x stays exactly W/2, angle exactly 0 (no torque at rudder 0)
    dt, damp = 0.1 * speed, 0.4 ** (0.1 * speed)
    v = 0.0
    for _ in ys:
        v = v * damp + 20.0 * dt                                    # f/m = 100/5
    assert env.pose[0, 4] == pytest.approx(v, rel=1e-12) and env.pose[0, 3] == 0.0
    assert v < 20.0 * dt / (1.0 - damp)                             # below the terminal velocity of the table


def test_rudder_then_thrust_known_answers():
    env = _env(10)
    out = env.step(np.array([[1], [0], [0], [0], [0]], dtype=np.int32))
    obs = out["obs"][:, 0, 16:]
    assert list(obs[:, 2]) == [-5.0] * 5                            # rudder -5 after the single action 1
    # a rudder action alone produces no force or torque: the pose of step 1 is the spawn pose
    assert obs[0, 0] == 300.0 and obs[0, 1] == 25.0 and obs[0, 3] == 0.0
    # thrust at local point (-rudder, 0): torque = -100 * rudder = +500 about a moment of 3087.5
    assert np.allclose(obs[1:, 3], [0.0, 0.161943, 0.388664, 0.641296], atol=1e-6)
    assert env.pose[0, 5] == pytest.approx(0.262996, abs=1e-6)
    assert env.pose[0, 0] < 300.0                                   # positive angle = counter-clockwise: drift toward -x
    assert env.moment == pytest.approx(3087.5, rel=1e-12)
    assert cbind.moment_for_poly(5.0, env.ship_hull()) == pytest.approx(3087.5, rel=1e-12)


def test_noop_from_rest_and_spawn_observation():
    env = _env(10)
    obs0 = env.hist.copy()
    assert np.all(obs0[0, :16] == -1.0) and list(obs0[0, 16:20]) == [300.0, 25.0, 0.0, 0.0] and np.all(obs0[0, 22:] == -1.0)
    out = env.step(np.array([[1], [2]], dtype=np.int32))           # rudder -5, rudder back to 0: the hull never moves
    assert np.array_equal(env.pose[0], [300.0, 25.0, 0.0, 0.0, 0.0, 0.0])
    # at the spawn pose the lidar origin is (310, 47.5) and no ray reaches a bank of the default map: readings stay -1
    assert np.all(out["obs"][:, 0, 22:] == -1.0)
    assert np.all(out["reward"] == -0.01) and not out["done"].any()


def test_episode_length_cap_and_reward_priority():
    env = _env(1, max_steps=7)                                      # dt = 0.1: the ship barely moves in 7 steps
    out = env.step(np.full((7, 1), 1, dtype=np.int32))
    assert list(out["done"][:, 0]) == [0] * 6 + [1]                 # exactly MAX_STEPS steps (ship_env.py:131,152)
    assert np.all(out["reward"] == -0.01)


# ---------------------------------------------------------------------------------------------- geometry
def _random_convex(rng, n, cx, cy, r):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = r * rng.uniform(0.6, 1.0, n)
    return cbind.convex_hull(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1))


def _inside(h, p, eps=0.0):
    e = np.roll(h, -1, axis=0) - h
    return np.all(e[:, 0] * (p[1] - h[:, 1]) - e[:, 1] * (p[0] - h[:, 0]) >= -eps)      # CCW: left of every edge


def _clip(h, a, b):
    """Cyrus-Beck: parameter range of a + t (b - a) inside the convex CCW polygon h, or None."""
    t0, t1 = 0.0, 1.0
    d = b - a
    for i in range(len(h)):
        p, q = h[i], h[(i + 1) % len(h)]
        nrm = np.array([q[1] - p[1], -(q[0] - p[0])])               # outward normal of a CCW edge
        num, den = nrm @ (a - p), nrm @ d
        if abs(den) < 1e-300:
            if num > 0:
                return None
            continue
        t = -num / den
        if den < 0:
            t0 = max(t0, t)
        else:
            t1 = min(t1, t)
    return (t0, t1) if t0 <= t1 else None


def test_segment_query_matches_cyrus_beck_clipping():
    rng = np.random.RandomState(0)
    n_hit = n_inside = n_miss = 0
    for _ in range(3000):
        h = _random_convex(rng, rng.randint(3, 13), 0.0, 0.0, 100.0)
        a = rng.uniform(-250, 250, 2)
        ang = rng.uniform(0, 2 * np.pi)
        b = a + 100.0 * np.array([np.cos(ang), np.sin(ang)])
        hit, point, alpha, margin = cbind.segment_query(h, a, b)
        if margin < 1e-6:
            continue                                               # grazing: either answer is legitimate
        if _inside(h, a):
            # cpShapeSegmentQuery: start inside the shape => reported at alpha 0 with `point` left at the segment END
            assert hit and alpha == 0.0 and np.allclose(point, b)
            n_inside += 1
            continue
        rng_t = _clip(h, a, b)
        if rng_t is None:
            assert not hit
            n_miss += 1
        else:
            assert hit and alpha == pytest.approx(rng_t[0], abs=1e-9)
            assert np.allclose(point, a + rng_t[0] * (b - a), atol=1e-7)
            n_hit += 1
    assert min(n_hit, n_inside, n_miss) > 100


def _seg_intersect(p, q, r, s):
    def orient(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    o1, o2, o3, o4 = orient(p, q, r), orient(p, q, s), orient(r, s, p), orient(r, s, q)
    return (o1 > 0) != (o2 > 0) and (o3 > 0) != (o4 > 0)


def test_polygon_contact_matches_exhaustive_test():
    rng = np.random.RandomState(1)
    ship = np.array([[0, 0], [0, 30], [10, 45], [20, 30], [20, 0]], dtype=float)
    n_touch = n_apart = 0
    for _ in range(3000):
        bank = _random_convex(rng, rng.randint(3, 13), 0.0, 0.0, 120.0)
        th = rng.uniform(-np.pi, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        hull = cbind.convex_hull(ship @ R.T + rng.uniform(-200, 200, 2))
        touch, sep = cbind.polys_touch(hull, bank)
        if abs(sep) < 1e-6:
            continue
        brute = any(_inside(bank, v) for v in hull) or any(_inside(hull, v) for v in bank) or any(
            _seg_intersect(hull[i], hull[(i + 1) % len(hull)], bank[j], bank[(j + 1) % len(bank)])
            for i in range(len(hull)) for j in range(len(bank)))
        assert touch == brute
        n_touch += touch
        n_apart += not touch
    assert min(n_touch, n_apart) > 300


def test_point_distance_matches_edgewise_minimum():
    rng = np.random.RandomState(2)
    for _ in range(2000):
        h = _random_convex(rng, rng.randint(3, 13), 0.0, 0.0, 50.0)
        p = rng.uniform(-100, 100, 2)
        best = np.inf
        for i in range(len(h)):
            a, b = h[i], h[(i + 1) % len(h)]
            t = np.clip((p - a) @ (b - a) / ((b - a) @ (b - a)), 0.0, 1.0)
            best = min(best, np.linalg.norm(p - (a + t * (b - a))))
        want = -best if _inside(h, p) else best                     # cpPolyShapePointQuery: negative inside
        assert cbind.poly_point_distance(h, p) == pytest.approx(want, abs=1e-9)


def test_scenario_picks_are_uniform_and_independent_of_neighbours():
    """The scenario an env plays next is an integer hash of (seed, global env id, episode) (the reference has no bank: it
    builds a level per reset, game.py:271-272).  What the batched env needs of it: uniform over the bank, no correlation
    between consecutive episodes of an env or between neighbouring envs, sensitive to every key word."""
    n_scen, n_env, n_ep = 64, 512, 256
    picks = np.array([[cbind.pick_scenario(12345, g, ep, n_scen) for ep in range(n_ep)] for g in range(n_env)])
    assert picks.min() == 0 and picks.max() == n_scen - 1
    counts = np.bincount(picks.ravel(), minlength=n_scen)
    expect = picks.size / n_scen
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert chi2 < 120, chi2                     # 63 degrees of freedom: mean 63, sd 11
    # consecutive episodes of an env, neighbouring envs at the same episode: pairs uniform over n_scen^2 -> equal with p = 1/n_scen
    same_next = (picks[:, 1:] == picks[:, :-1]).mean()
    same_nb = (picks[1:] == picks[:-1]).mean()
    assert abs(same_next - 1 / n_scen) < 0.004 and abs(same_nb - 1 / n_scen) < 0.004, (same_next, same_nb)
    # every word of the key matters: high seed word, high env-id word
    a = np.array([cbind.pick_scenario(12345, g, 3, 1 << 20) for g in range(64)])
    b = np.array([cbind.pick_scenario(12345 + (1 << 32), g, 3, 1 << 20) for g in range(64)])
    c = np.array([cbind.pick_scenario(12345, g + (1 << 32), 3, 1 << 20) for g in range(64)])
    assert (a != b).mean() > 0.95 and (a != c).mean() > 0.95
