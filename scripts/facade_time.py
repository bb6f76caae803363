"""Exploration: steps/s of the single-env drop-in facade (BASELINE.json configs[0] and E = 1 hd)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import numpy as np
import formation_gym

for scen, n in (("basic_formation_env", 3), ("formation_hd_env", 9), ("formation_hd_env", 27), ("formation_hd_obs_env", 4)):
    env = formation_gym.make_env(scen, False, n, 25)
    env.reset()
    def one():
        act = [sp.sample() for sp in env.action_space]
        o, r, d, i = env.step(act)
        if np.all(d):
            env.reset()
    for _ in range(30): one()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < 2.0:
        one(); k += 1
    dt = time.perf_counter() - t0
    print("%s N=%d: %.0f env-steps/s (%.1f us/step)" % (scen, n, k / dt, dt / k * 1e6), flush=True)
