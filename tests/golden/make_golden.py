"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It drives jc-bao/gym-formation's own ``formation_gym.make_env(...).step()`` through
``oracle/ref_harness.py`` (in-process stubs for imp/gym/multiagent, reference files untouched)
from injected states and freezes inputs + outputs as float64 ``.npz`` files.  The reference has
no tests or golden vectors of its own, so these files are the parity pins:
``tests/test_oracle_golden.py`` checks the oracle against them (CPU), ``tests/test_gpu_parity.py``
checks the CUDA path against them (GPU).

All injected inputs are rounded to float32-representable values so that the fp32 CUDA build,
the fp64 CUDA build, the oracle and the reference all start from bit-identical states (Q15).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402

OBS_ROWS_243 = [0, 1, 121, 242]


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def sample_state(rng, n, spread, n_landmarks=None):
    L = n if n_landmarks is None else n_landmarks
    pos = f32(rng.uniform(-spread, spread, (n, 2)))
    vel = f32(rng.uniform(-0.5, 0.5, (n, 2)))
    act = f32(rng.uniform(-1, 1, (n, 2)))
    lm = f32(rng.uniform(-1, 1, (L, 2)))
    shape = f32(lm - lm.mean(0))
    ivel = f32(rng.uniform(-1, 1, 2))
    return pos, vel, act, lm, shape, ivel


def hd_single(n, spread, samples, seed, episode_length=25, obs_rows=None):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("formation_hd_env", n, episode_length)
    out = {k: [] for k in ("pos0 vel0 act lm0 shape ivel step0 pos vel obs reward indiv done "
                           "landmarks").split()}
    for s in range(samples):
        pos, vel, act, lm, shape, ivel = sample_state(rng, n, spread)
        # last sample of every file sits one step before the episode end -> done=True
        step0 = episode_length - 1 if s == samples - 1 else int(rng.integers(0, episode_length - 1))
        rh.inject_state(env, pos, vel, shape, ivel, lm, step0)
        r = rh.reference_step(env, act)
        obs = r["obs"] if obs_rows is None else r["obs"][obs_rows]
        for k, v in (("pos0", pos), ("vel0", vel), ("act", act), ("lm0", lm), ("shape", shape),
                     ("ivel", ivel), ("step0", step0), ("pos", r["pos"]), ("vel", r["vel"]),
                     ("obs", obs), ("reward", r["reward"]), ("indiv", r["indiv"]),
                     ("done", r["done"]), ("landmarks", r["landmarks"])):
            out[k].append(v)
    res = {k: np.stack(v) for k, v in out.items()}
    if obs_rows is not None:
        res["obs_rows"] = np.array(obs_rows)
    return res


def hd_traj(n, spread, steps, seed, episode_length=100, obs_rows=None):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("formation_hd_env", n, episode_length)
    pos, vel, _, lm, shape, ivel = sample_state(rng, n, spread)
    vel = np.zeros_like(vel)
    acts = f32(rng.uniform(-1, 1, (steps, n, 2)))
    rh.inject_state(env, pos, vel, shape, ivel, lm, 0)
    P, V, R, I, O = [], [], [], [], None
    for t in range(steps):
        r = rh.reference_step(env, acts[t])
        P.append(r["pos"]); V.append(r["vel"]); R.append(r["reward"]); I.append(r["indiv"])
        O = r["obs"]
    res = dict(pos0=pos, vel0=vel, lm0=lm, shape=shape, ivel=ivel, acts=acts,
               pos=np.stack(P), vel=np.stack(V), reward=np.stack(R), indiv=np.stack(I),
               obs_last=O if obs_rows is None else O[obs_rows])
    if obs_rows is not None:
        res["obs_rows"] = np.array(obs_rows)
    return res


def basic_single(n, spread, samples, seed, tweak=None, walls=None, episode_length=25):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("basic_formation_env", n, episode_length)
    if tweak is not None:
        for a, (m, acc, vmax) in zip(env.world.agents, tweak):
            a.initial_mass, a.accel, a.max_speed = m, acc, vmax
    if walls is not None:
        core = sys.modules["formation_gym.core"]
        env.world.walls = [core.Wall(o, ap, (e0, e1), w, True) for (o, ap, e0, e1, w) in walls]
    L = len(env.world.landmarks)
    out = {k: [] for k in "pos0 vel0 act lm0 step0 pos vel obs reward indiv done".split()}
    for s in range(samples):
        pos, vel, act, lm, _, _ = sample_state(rng, n, spread, L)
        step0 = episode_length - 1 if s == samples - 1 else int(rng.integers(0, episode_length - 1))
        rh.inject_state(env, pos, vel, None, None, lm, step0)
        r = rh.reference_step(env, act)
        for k, v in (("pos0", pos), ("vel0", vel), ("act", act), ("lm0", lm), ("step0", step0),
                     ("pos", r["pos"]), ("vel", r["vel"]), ("obs", r["obs"]),
                     ("reward", r["reward"]), ("indiv", r["indiv"]), ("done", r["done"])):
            out[k].append(v)
    res = {k: np.stack(v) for k, v in out.items()}
    if tweak is not None:
        res["tweak"] = np.array([[m, -1.0 if acc is None else acc, -1.0 if vm is None else vm]
                                 for (m, acc, vm) in tweak], np.float64)
    if walls is not None:
        res["walls"] = np.array([[0.0 if o == 'H' else 1.0, ap, e0, e1, w]
                                 for (o, ap, e0, e1, w) in walls], np.float64)
    return res


def basic_traj(n, spread, steps, seed, episode_length=50):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("basic_formation_env", n, episode_length)
    L = len(env.world.landmarks)
    pos, vel, _, lm, _, _ = sample_state(rng, n, spread, L)
    vel = np.zeros_like(vel)
    acts = f32(rng.uniform(-1, 1, (steps, n, 2)))
    rh.inject_state(env, pos, vel, None, None, lm, 0)
    P, V, R, I, O = [], [], [], [], None
    for t in range(steps):
        r = rh.reference_step(env, acts[t])
        P.append(r["pos"]); V.append(r["vel"]); R.append(r["reward"]); I.append(r["indiv"])
        O = r["obs"]
    return dict(pos0=pos, vel0=vel, lm0=lm, acts=acts, pos=np.stack(P), vel=np.stack(V),
                reward=np.stack(R), indiv=np.stack(I), obs_last=O)


def hd_hetero(seed):
    """hd N=9 with per-agent mass / accel / max_speed set (exercises core.py:235-236,271-276,
    314-318 and the accel-twice quirk Q20)."""
    n = 9
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env("formation_hd_env", n, 25)
    tweak = [(1.0 + 0.25 * (i % 3), 3.0 + 0.5 * i, 0.5 + 0.125 * i) for i in range(n)]
    for a, (m, acc, vmax) in zip(env.world.agents, tweak):
        a.initial_mass, a.accel, a.max_speed = m, acc, vmax
    out = {k: [] for k in ("pos0 vel0 act lm0 shape ivel step0 pos vel obs reward indiv done"
                           ).split()}
    for s in range(6):
        pos, vel, act, lm, shape, ivel = sample_state(rng, n, 0.12)
        rh.inject_state(env, pos, vel, shape, ivel, lm, 3)
        r = rh.reference_step(env, act)
        for k, v in (("pos0", pos), ("vel0", vel), ("act", act), ("lm0", lm), ("shape", shape),
                     ("ivel", ivel), ("step0", 3), ("pos", r["pos"]), ("vel", r["vel"]),
                     ("obs", r["obs"]), ("reward", r["reward"]), ("indiv", r["indiv"]),
                     ("done", r["done"])):
            out[k].append(v)
    res = {k: np.stack(v) for k, v in out.items()}
    res["tweak"] = np.array(tweak, np.float64)
    return res


def hd_nan_quirk():
    """Coincident agents -> 0/0 -> NaN (core.py:312; train/README.md:194-197)."""
    n = 3
    env = rh.make_reference_env("formation_hd_env", n, 25)
    pos = f32([[0.25, 0.5], [0.25, 0.5], [-0.5, 0.125]])
    vel = np.zeros((n, 2))
    lm = f32([[0.0, 1.0], [1.0, 0.0], [-1.0, 0.0]])
    shape = f32(lm - lm.mean(0))
    ivel = f32([0.5, -0.25])
    act = f32([[0.5, 0.5], [-0.5, 0.25], [1.0, -1.0]])
    rh.inject_state(env, pos, vel, shape, ivel, lm, 0)
    with np.errstate(all="ignore"):
        r = rh.reference_step(env, act)
    return dict(pos0=pos, vel0=vel, act=act, lm0=lm, shape=shape, ivel=ivel,
                pos=r["pos"], vel=r["vel"], obs=r["obs"], indiv=r["indiv"], reward=r["reward"])


def api_contract():
    """Types / shapes / lengths of the reference API (environment.py:113-156) as JSON."""
    res = {}
    for scen, n in (("formation_hd_env", 9), ("basic_formation_env", 3)):
        np.random.seed(0)
        env = rh.make_reference_env(scen, n)
        obs_n = env.reset()
        act_n = [np.zeros(2) for _ in range(n)]
        o, r, d, i = env.step(act_n)
        res[scen] = dict(
            num_agents=env.num_agents, world_length=env.world_length,
            shared_reward=bool(env.shared_reward),
            obs_len=len(obs_n), obs_dim=int(obs_n[0].shape[0]), obs_dtype=str(obs_n[0].dtype),
            action_shape=list(env.action_space[0].shape),
            action_low=float(env.action_space[0].low), action_high=float(env.action_space[0].high),
            obs_space_shape=list(env.observation_space[0].shape),
            share_obs_space_shape=list(env.share_observation_space[0].shape),
            reward_type=type(r).__name__, reward_inner_type=type(r[0]).__name__,
            reward_inner_len=len(r[0]), reward_aliased=bool(r[0] is r[1]),
            done_type=type(d[0]).__name__, info_keys=sorted(i[0].keys()),
            dim_c=int(env.world.dim_c), agent_size=float(env.world.agents[0].size),
            n_landmarks=len(env.world.landmarks),
        )
    return res


def main():
    import scipy
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0],
                generator="tests/golden/make_golden.py",
                reference="jc-bao/gym-formation @ /root/reference (unmodified, via oracle/ref_harness.py)")
    save = lambda name, d: np.savez_compressed(os.path.join(HERE, name), **d)  # noqa: E731

    for n, spread_far, spread_near, samples in ((3, 1.0, 0.05, 8), (9, 1.0, 0.12, 8),
                                                (27, 1.0, 0.25, 8)):
        save("hd_n%d_spread.npz" % n, hd_single(n, spread_far, samples, 1000 + n))
        save("hd_n%d_clustered.npz" % n, hd_single(n, spread_near, samples, 2000 + n))
        print("hd single", n, flush=True)
    save("hd_n243_spread.npz", hd_single(243, 1.0, 2, 1243, obs_rows=OBS_ROWS_243))
    save("hd_n243_clustered.npz", hd_single(243, 0.6, 2, 2243, obs_rows=OBS_ROWS_243))
    print("hd single 243", flush=True)
    for n, spread in ((3, 0.06), (9, 0.15), (27, 0.3)):
        save("hd_n%d_traj25.npz" % n, hd_traj(n, spread, 25, 3000 + n))
    save("hd_n243_traj25.npz", hd_traj(243, 0.7, 25, 3243, obs_rows=OBS_ROWS_243))
    print("hd traj", flush=True)
    save("hd_n9_hetero.npz", hd_hetero(4009))
    save("hd_n3_nan.npz", hd_nan_quirk())
    save("basic_n3_spread.npz", basic_single(3, 1.0, 8, 5003))
    save("basic_n3_clustered.npz", basic_single(3, 0.2, 8, 6003))
    save("basic_n5_clustered.npz", basic_single(5, 0.3, 6, 6005))
    save("basic_n3_hetero.npz", basic_single(
        3, 0.25, 6, 7003, tweak=[(1.0, None, 0.5), (2.0, 4.0, None), (0.5, 3.0, 0.75)]))
    save("basic_n3_walls.npz", basic_single(
        3, 0.6, 8, 8003, walls=[('H', 0.5, -0.4, 0.4, 0.1), ('V', -0.3, -1.0, 1.0, 0.2)]))
    save("basic_n3_traj25.npz", basic_traj(3, 0.25, 25, 9003))
    print("basic", flush=True)
    with open(os.path.join(HERE, "api_contract.json"), "w") as f:
        json.dump(dict(meta=meta, api=api_contract()), f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "META.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
