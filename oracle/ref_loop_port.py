"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- per-env, per-object Python loop port of the
reference's step path, used as the timed CPU baseline ("the reference's per-env numpy loop").

The unmodified reference cannot travel to the GPU box (no /root/reference there), so the CPU
baseline that bench.py reports is this port.  Unlike oracle/mpe_oracle.py (batched, vectorised,
built for checking) it keeps the reference's WORK STRUCTURE so that its speed is representative:
one Python object per agent holding tiny float64 arrays, a Python double loop over entity pairs
(formation_gym/core.py:240-262,289-322), per-agent observation built with np.append/concatenate
(formation_gym/envs/formation_hd_env.py:38-59) and the reward evaluated 2N times per step
(formation_gym/environment.py:128,130), each evaluation running two directed Hausdorff distances
(formation_hd_env.py:66; scipy's function when scipy is importable -- the same third-party call the
reference makes -- else the brute-force restatement).  tests/test_oracle_golden.py checks it
against the golden fixtures; DESIGN.md records its speed next to the unmodified reference's,
measured in the build container.
"""
import os
import time

import numpy as np

try:
    from scipy.spatial.distance import directed_hausdorff as _dh
    HAVE_SCIPY = True
except Exception:  # pragma: no cover
    HAVE_SCIPY = False

    def _dh(u, v):
        d = np.sqrt(((u[:, None, :] - v[None, :, :]) ** 2).sum(2))
        return (d.min(1).max(), 0, 0)


class _Ent(object):
    __slots__ = ("pos", "vel", "c", "u", "size", "mass", "movable", "collide")


class RefLoopEnv(object):
    """formation_hd_env / basic_formation_env for ONE env, reference-style loops."""

    def __init__(self, scenario="formation_hd_env", num_agents=9, episode_length=25, num_landmarks=3):
        self.hd = scenario == "formation_hd_env"
        self.n = num_agents
        self.world_length = episode_length
        self.dt, self.damping, self.cf, self.margin = 0.1, 0.25, 1e2, 1e-3
        self.agents, self.landmarks = [], []
        for _ in range(num_agents):
            a = _Ent()
            a.size, a.mass, a.movable, a.collide = (0.03 if self.hd else 0.1), 1.0, True, True
            self.agents.append(a)
        for _ in range(num_agents if self.hd else num_landmarks):
            l = _Ent()
            l.size, l.mass, l.movable, l.collide = (0.01 if self.hd else 0.05), 1.0, False, False
            self.landmarks.append(l)
        self.current_step = 0
        self.reset()

    # -- Scenario.reset_world
    def reset(self):
        self.current_step = 0
        for a in self.agents:
            a.pos = np.random.uniform(-1, +1, 2)
            a.vel = np.zeros(2)
            a.c = np.zeros(2)
        raw = []
        for l in self.landmarks:
            l.pos = np.random.uniform(-1, +1, 2)
            l.vel = np.zeros(2)
            raw.append(l.pos)
        if self.hd:
            self.ideal_shape = raw - np.mean(raw, 0)
            self.ideal_vel = np.random.uniform(-1, +1, 2)
        return [self._obs(a) for a in self.agents]

    # -- World.step
    def _pair_force(self, ea, eb):
        if (not ea.collide) or (not eb.collide):
            return None, None
        if (not ea.movable) and (not eb.movable):
            return None, None
        delta = ea.pos - eb.pos
        dist = np.linalg.norm(delta)
        dist_min = ea.size + eb.size
        k = self.margin
        pen = np.logaddexp(0, -(dist - dist_min) / k) * k
        force = self.cf * delta / dist * pen
        ratio = eb.mass / ea.mass
        return ratio * force, -(1 / ratio) * force

    def _world_step(self):
        ents = self.agents + self.landmarks
        F = [None] * len(ents)
        for i, a in enumerate(self.agents):
            F[i] = a.mass * a.u + 0.0
        for ia, ea in enumerate(ents):
            for ib in range(ia + 1, len(ents)):
                fa, fb = self._pair_force(ea, ents[ib])
                if fa is not None:
                    F[ia] = fa + (0.0 if F[ia] is None else F[ia])
                if fb is not None:
                    F[ib] = fb + (0.0 if F[ib] is None else F[ib])
        for i, e in enumerate(ents):
            if not e.movable:
                continue
            e.vel = e.vel * (1 - self.damping)
            if F[i] is not None:
                e.vel += (F[i] / e.mass) * self.dt
            e.pos += e.vel * self.dt
        for a in self.agents:
            a.c = np.zeros(2)

    # -- Scenario.observation / reward
    def _obs(self, agent):
        if self.hd:
            u = [a.pos for a in self.agents]
            v = [l.pos for l in self.landmarks]
            delta = np.mean(u, 0) - np.mean(v, 0)
            for l in self.landmarks:
                l.pos = l.pos + delta
            other_pos, comm = np.array([]), np.array([])
            for o in self.agents:
                if o is agent:
                    continue
                comm = np.append(comm, o.c)
                other_pos = np.append(other_pos, o.pos - agent.pos)
            return np.concatenate((agent.vel, other_pos, comm, self.ideal_shape.flatten(), self.ideal_vel))
        ent = [l.pos - agent.pos for l in self.landmarks]
        other_pos, comm = [], []
        for o in self.agents:
            if o is agent:
                continue
            comm.append(o.c)
            other_pos.append(o.pos - agent.pos)
        return np.concatenate([agent.vel] + [agent.pos] + ent + other_pos + comm)

    def _reward(self, agent):
        if self.hd:
            shape = [a.pos for a in self.agents]
            shape = shape - np.mean(shape, 0)
            rew = -max(_dh(shape, self.ideal_shape)[0], _dh(self.ideal_shape, shape)[0])
            mean_vel = np.mean([a.vel for a in self.agents], axis=0)
            rew -= np.linalg.norm(self.ideal_vel - mean_vel)
            for a in self.agents:
                if a is not agent and np.linalg.norm(a.pos - agent.pos) < (a.size + agent.size) / 2:
                    rew -= 1
            return rew
        rew = 0
        for l in self.landmarks:
            rew -= min(np.linalg.norm(a.pos - l.pos) for a in self.agents)
        for a in self.agents:
            if np.linalg.norm(a.pos - agent.pos) < (a.size + agent.size):
                rew -= 1
        return rew

    # -- MultiAgentEnv.step
    def step(self, action_n):
        self.current_step += 1
        for a, act in zip(self.agents, action_n):
            a.u = np.array(act, dtype=np.float64) * 5.0
        self._world_step()
        obs_n, reward_n, done_n, info_n = [], [], [], []
        for a in self.agents:
            obs_n.append(self._obs(a))
            reward_n.append([self._reward(a)])
            done_n.append(self.current_step >= self.world_length)
            info_n.append({'individual_reward': self._reward(a)})     # evaluated a second time
        reward = np.sum(reward_n)
        reward_n = [[reward]] * self.n
        return obs_n, reward_n, done_n, info_n


def _worker(args):
    scenario, n, episode_length, seconds, seed = args
    np.random.seed(seed)
    env = RefLoopEnv(scenario, n, episode_length)
    acts = lambda: [np.random.uniform(-1, 1, 2) for _ in range(n)]  # noqa: E731
    for _ in range(2 if n > 100 else min(episode_length, 10)):      # warm-up
        env.step(acts())
    steps, t0 = 0, time.perf_counter()
    while True:
        _, _, done_n, _ = env.step(acts())
        steps += 1
        if all(done_n):
            env.reset()
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return steps, dt


def time_port(scenario="formation_hd_env", num_agents=9, episode_length=25, seconds=5.0, procs=None):
    """Random-policy stepping rate of the port on `procs` forked processes, one env each (the way
    the reference parallelises: train/maddpg-v2/utils/env_wrappers.py:48-55).
    Returns dict(agent_steps_per_s, env_steps_per_s, procs, env_steps, seconds)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    args = [(scenario, num_agents, episode_length, seconds, 1234 + k) for k in range(procs)]
    if procs == 1:
        res = [_worker(args[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_worker, args)
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return dict(agent_steps_per_s=steps * num_agents / wall, env_steps_per_s=steps / wall,
                procs=procs, env_steps=steps, seconds=wall, scipy=HAVE_SCIPY)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="formation_hd_env")
    ap.add_argument("--agents", type=int, default=9)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--procs", type=int, default=1)
    a = ap.parse_args()
    print(time_port(a.scenario, a.agents, 25, a.seconds, a.procs))
