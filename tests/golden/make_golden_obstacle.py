"""Golden fixtures for formation_hd_obs_env (movable colliding obstacle landmarks) from the
UNMODIFIED reference (build container only):  python tests/golden/make_golden_obstacle.py

  obstacle_n*.npz       single env.step from injected states (agents, goal landmarks, obstacles and their
                        velocities), obstacles placed among the agents so that agent-obstacle and
                        obstacle-obstacle contacts occur
  obstacle_traj50.npz   reset (seeded) + 50 steps of the stock scenario (4 agents, 4 goals, 3 obstacles that fall
                        through the agents), with the raw U(0,1) draws of the reset so that the reset restatement
                        can be pinned too
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402

SCN = "formation_hd_obs_env"


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def inject(env, pos, vel, goals, obst, obst_vel, step):
    w = env.world
    L = len(goals)
    for i, a in enumerate(w.agents):
        a.state.p_pos = np.array(pos[i], np.float64)
        a.state.p_vel = np.array(vel[i], np.float64)
        a.state.c = np.zeros(w.dim_c)
    for k, l in enumerate(w.landmarks):
        if k < L:
            l.state.p_pos = np.array(goals[k], np.float64)
            l.state.p_vel = np.zeros(w.dim_p)
        else:
            l.state.p_pos = np.array(obst[k - L], np.float64)
            l.state.p_vel = np.array(obst_vel[k - L], np.float64)
    env.current_step = int(step)


def read(env, L):
    w = env.world
    lm = np.stack([l.state.p_pos for l in w.landmarks]).astype(np.float64)
    lv = np.stack([np.asarray(l.state.p_vel, np.float64) for l in w.landmarks])
    return lm[:L], lm[L:], lv[L:]


def step(env, act, L):
    r = rh.reference_step(env, act)
    goals, obst, ov = read(env, L)
    r.update(goals=goals, obst=obst, obst_vel=ov)
    return r


def make_single(n, samples, seed, spread):
    rng = np.random.default_rng(seed)
    env = rh.make_reference_env(SCN, n, 50)
    sc = rh.scenario_of(env)
    L, O = sc.num_landmarks, sc.num_obstacles
    keys = "pos0 vel0 act goals obst0 obst_vel0 step0 pos vel obst obst_vel obs reward indiv done".split()
    out = {k: [] for k in keys}
    for s in range(samples):
        pos = f32(rng.uniform(-spread, spread, (n, 2)))
        vel = f32(rng.uniform(-0.5, 0.5, (n, 2)))
        act = f32(rng.uniform(-1, 1, (n, 2)))
        goals = f32(rng.uniform(-1, 1, (L, 2)))
        obst = f32(rng.uniform(-spread - 0.2, spread + 0.2, (O, 2)))
        if s % 4 == 3:
            obst[:, 1] -= 2.4            # some below the floor: the velocity rule's other branch
        ov = f32(rng.uniform(-1, 1, (O, 2))) if s % 2 else np.tile([0.0, -1.0], (O, 1))
        step0 = 49 if s == samples - 1 else int(rng.integers(0, 49))
        inject(env, pos, vel, goals, obst, ov, step0)
        r = step(env, act, L)
        for k, v in (("pos0", pos), ("vel0", vel), ("act", act), ("goals", goals), ("obst0", obst), ("obst_vel0", ov),
                     ("step0", step0), ("pos", r["pos"]), ("vel", r["vel"]), ("obst", r["obst"]),
                     ("obst_vel", r["obst_vel"]), ("obs", r["obs"]), ("reward", r["reward"]), ("indiv", r["indiv"]),
                     ("done", r["done"])):
            out[k].append(v)
        assert np.array_equal(r["goals"], goals)
    return {k: np.stack(v) for k, v in out.items()}


def make_traj(seed=7, steps=50):
    env = rh.make_reference_env(SCN, 4, 50)
    sc = rh.scenario_of(env)
    n, L, O = 4, sc.num_landmarks, sc.num_obstacles
    np.random.seed(seed)
    st = np.random.get_state()
    obs0 = np.stack(env.reset())
    np.random.set_state(st)
    raw = np.random.random_sample((n + L + O, 2))            # the same U(0,1) stream the reset consumed
    pos0 = np.stack([a.state.p_pos for a in env.world.agents])
    goals, obst0, ov0 = read(env, L)
    rng = np.random.default_rng(seed)
    # steer the agents up into the falling obstacles so that contacts happen
    acts = f32(np.clip(rng.uniform(-1, 1, (steps, n, 2)) + np.array([0.0, 0.6]), -1, 1))
    tr = {k: [] for k in "pos vel obst obst_vel obs reward indiv done".split()}
    for t in range(steps):
        r = step(env, acts[t], L)
        for k in tr:
            tr[k].append(r[k])
    d = {k: np.stack(v) for k, v in tr.items()}
    d.update(raw=raw, pos0=pos0, goals=goals, obst0=obst0, obst_vel0=ov0, obs0=obs0, act=acts)
    return d


if __name__ == "__main__":
    for n, spread, seed in ((4, 0.5, 1), (3, 0.3, 2), (9, 0.6, 3), (27, 1.0, 4), (2, 0.2, 5)):
        d = make_single(n, 16, 900 + seed, spread)
        np.savez_compressed(os.path.join(HERE, "obstacle_n%d.npz" % n), **d)
        print("obstacle", n, d["obs"].shape, float(np.abs(d["vel"] - 0.75 * d["vel0"] - 0.5 * d["act"]).max()))
    d = make_traj()
    np.savez_compressed(os.path.join(HERE, "obstacle_traj50.npz"), **d)
    print("traj", d["obs"].shape, "min agent-obstacle distance",
          float(np.sqrt(((d["pos"][:, :, None] - d["obst"][:, None]) ** 2).sum(-1)).min()))
