"""Shared helpers for the test-suite (golden loading, bank packing)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_trajectories():
    z = np.load(os.path.join(GOLDEN, "trajectories.npz"))
    n = int(z["n_episodes"])
    eps = []
    for i in range(n):
        pre = "e%d_" % i
        eps.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    return eps


def pack_bank(hulls_list, goals_list, maxv=None):
    """hulls_list: [(hull0 (m0,2), hull1 (m1,2)), ...]; goals_list: [(5,2), ...] -> float64 bank dict."""
    S = len(hulls_list)
    if maxv is None:
        maxv = max(max(len(h0), len(h1)) for h0, h1 in hulls_list)
    hull_xy = np.zeros((S, 2, maxv, 2))
    hull_n = np.zeros((S, 2), dtype=np.int32)
    goals = np.zeros((S, 5, 2))
    for s, ((h0, h1), g) in enumerate(zip(hulls_list, goals_list)):
        for b, h in enumerate((h0, h1)):
            hull_xy[s, b, :len(h)] = h
            hull_n[s, b] = len(h)
        goals[s] = g
    return dict(hull_xy=hull_xy, hull_n=hull_n, goals=goals)
