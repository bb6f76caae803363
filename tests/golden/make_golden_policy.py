"""Golden fixtures for the hand-written controller, from the UNMODIFIED reference:
``formation_gym.get_action_BFS(formation_gym.ezpolicy, obs_n, 3)`` (formation_gym/__init__.py:19-98) on
observations produced by the reference env from injected states.

    python tests/golden/make_golden_policy.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def make(n_agents, samples, seed, spread=1.0, near=False):
    fg = rh.load_reference()
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("formation_hd_env", n_agents, 25)
    out = {k: [] for k in ("pos", "vel", "shape", "ivel", "act")}
    for s in range(samples):
        lm = f32(rng.uniform(-1, 1, (n_agents, 2)))
        shape = f32(lm - lm.mean(0))
        if near:      # agents close to a translated copy of the target shape: exercises the `done` branch
            pos = f32(shape + rng.uniform(-1, 1, 2) + rng.normal(0, 1e-3 if s % 2 else 0.05, (n_agents, 2)))
        else:
            pos = f32(rng.uniform(-spread, spread, (n_agents, 2)))
        vel = f32(rng.uniform(-0.5, 0.5, (n_agents, 2)))
        ivel = f32(rng.uniform(-1, 1, 2))
        rh.inject_state(env, pos, vel, shape, ivel, lm, 0)
        obs_n = [env._get_obs(a) for a in env.agents]                 # scenario.observation per agent
        act = fg.get_action_BFS(fg.ezpolicy, obs_n, 3)                # THE reference controller
        for k, v in (("pos", pos), ("vel", vel), ("shape", shape), ("ivel", ivel), ("act", np.stack(act))):
            out[k].append(v)
    return {k: np.stack(v) for k, v in out.items()}


if __name__ == "__main__":
    for n, samples in ((3, 64), (9, 48), (27, 24), (81, 6)):
        d = make(n, samples, 900 + n)
        np.savez_compressed(os.path.join(HERE, "policy_bfs_n%d.npz" % n), **d)
        print("policy_bfs_n%d" % n, d["act"].shape)
    d = make(9, 24, 77, near=True)
    np.savez_compressed(os.path.join(HERE, "policy_bfs_n9_near.npz"), **d)
    d = make(3, 32, 78, near=True)
    np.savez_compressed(os.path.join(HERE, "policy_bfs_n3_near.npz"), **d)
    print("done")
