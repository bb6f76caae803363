"""Golden fixtures for formation_hd_partial_env / formation_hd_partial_range_env from
the UNMODIFIED reference (build container only):  python tests/golden/make_golden_partial.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def make(scenario, n, samples, seed, spread):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env(scenario, n, 25)
    sc = rh.scenario_of(env)
    L = len(env.world.landmarks)
    keys = "pos0 vel0 act lm step0 pos vel obs reward indiv done".split()
    out = {k: [] for k in keys}
    for s in range(samples):
        pos = f32(rng.uniform(-spread, spread, (n, 2)))
        vel = f32(rng.uniform(-0.5, 0.5, (n, 2)))
        act = f32(rng.uniform(-1, 1, (n, 2)))
        lm = f32(rng.uniform(-1, 1, (L, 2)))
        step0 = 24 if s == samples - 1 else int(rng.integers(0, 24))
        rh.inject_state(env, pos, vel, None, None, lm, step0)
        r = rh.reference_step(env, act)
        for k, v in (("pos0", pos), ("vel0", vel), ("act", act), ("lm", lm), ("step0", step0), ("pos", r["pos"]),
                     ("vel", r["vel"]), ("obs", r["obs"]), ("reward", r["reward"]), ("indiv", r["indiv"]),
                     ("done", r["done"])):
            out[k].append(v)
    d = {k: np.stack(v) for k, v in out.items()}
    d["num_obs"] = np.int64(getattr(sc, "num_obs", -1))
    d["obs_range"] = np.float64(getattr(sc, "obs_range", -1.0))
    return d


if __name__ == "__main__":
    for scenario, tag in (("formation_hd_partial_env", "partial"), ("formation_hd_partial_range_env", "range")):
        for n, spread, seed in ((5, 1.0, 1), (4, 0.15, 2), (9, 0.3, 3), (27, 0.5, 4), (2, 0.1, 5)):
            d = make(scenario, n, 16, 500 + seed + (0 if tag == "partial" else 50), spread)
            np.savez_compressed(os.path.join(HERE, "%s_n%d.npz" % (tag, n)), **d)
            print(tag, n, d["obs"].shape, int(d["num_obs"]), float(d["obs_range"]))
