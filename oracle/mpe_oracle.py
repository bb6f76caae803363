"""TEST INFRASTRUCTURE ONLY -- CPU oracle for gym-formation's MPE step path (numpy, float64).

Closed-form restatement of the reference algorithm (jc-bao/gym-formation, all citations
relative to /root/reference):

  * ``MultiAgentEnv.step / _set_action / _get_done``   formation_gym/environment.py:113-142,172-236
  * ``World.step`` and helpers                          formation_gym/core.py:206-322 (+325-362 walls)
  * hd scenario ``observation/reward/reset_world``      formation_gym/envs/formation_hd_env.py:38-95,119-121
  * basic scenario ``observation/reward/reset_world``   formation_gym/envs/basic_formation_env.py:29-65,89-91
  * partial-observation scenarios                       formation_gym/envs/formation_hd_partial_env.py:41-125,
                                                        formation_gym/envs/formation_hd_partial_range_env.py:41-113
  * obstacle scenario (movable colliding landmarks)     formation_gym/envs/formation_hd_obs_env.py:14-149

Third-party arithmetic restated here (the reference pins no versions --
``setup.py:4-17`` has no install_requires; this container has numpy 2.3.5 / scipy 1.18.1):
  * ``scipy.spatial.distance.directed_hausdorff(u, v)[0]`` == sqrt(max_i min_j |u_i - v_j|^2)
    (call site formation_hd_env.py:66) -- restated as a brute-force max-min.
  * ``np.logaddexp(0, t)`` (core.py:310) == stable softplus ``t > 0 ? t + log1p(exp(-t)) : log1p(exp(t))``;
    numpy's own ufunc is called here (the CUDA kernels spell it out, fg_kernels.cuh contact_force).

PINNING: the reference ships no tests, golden vectors or fixtures for this path,
so this oracle is pinned against OUTPUTS OF THE UNMODIFIED REFERENCE RUN IN THE BUILD CONTAINER:
``tests/golden/make_golden.py`` (committed) drives /root/reference through oracle/ref_harness.py
and freezes (state, action) -> (pos, vel, obs, rewards, done) into ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this module, and only as the checker / the timed CPU baseline.  The product
package must never import it.

All arrays are batched: pos/vel/act ``[E,N,2]``, ideal_shape ``[E,N,2]``, ideal_vel ``[E,2]``,
landmarks ``[E,L,2]``.  Accumulation order inside an env follows the reference (pairs a<b in
entity order, core.py:242-254) so float64 results are reproducible to the last bit or two.
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


@dataclass
class WorldParams:
    """World/Agent constants (core.py:45-110,112-139) as one flat record."""
    dt: float = 0.1                  # core.py:125
    damping: float = 0.25            # core.py:127
    contact_force: float = 1e2       # core.py:129
    contact_margin: float = 1e-3     # core.py:130
    sensitivity: float = 5.0         # environment.py:218
    agent_size: float = 0.03         # formation_hd_env.py:26 (basic: 0.1, basic_formation_env.py:18)
    mass: float = 1.0                # core.py:69,73-75
    accel: Optional[float] = None    # core.py:65   (scales the action twice when set: Q20)
    max_speed: Optional[float] = None  # core.py:64
    u_noise: Optional[float] = None  # core.py:97
    collide: bool = True             # formation_hd_env.py:24
    world_length: int = 100          # formation_hd_env.py:13,16 (basic: 50, core.py:113)
    # optional per-agent overrides [N] (heterogeneous mass/size/accel/max_speed)
    agent_mass: Optional[Sequence[float]] = None
    agent_sizes: Optional[Sequence[float]] = None
    agent_accel: Optional[Sequence[float]] = None
    agent_max_speed: Optional[Sequence[float]] = None
    walls: list = field(default_factory=list)   # (orient 'H'|'V', axis_pos, end0, end1, width, hard)

    def per_agent(self, n):
        mass = np.full(n, self.mass, np.float64) if self.agent_mass is None \
            else np.asarray(self.agent_mass, np.float64)
        size = np.full(n, self.agent_size, np.float64) if self.agent_sizes is None \
            else np.asarray(self.agent_sizes, np.float64)
        if self.agent_accel is not None:
            accel = np.asarray(self.agent_accel, np.float64)
        elif self.accel is not None:
            accel = np.full(n, self.accel, np.float64)
        else:
            accel = None
        if self.agent_max_speed is not None:
            vmax = np.asarray(self.agent_max_speed, np.float64)
        elif self.max_speed is not None:
            vmax = np.full(n, self.max_speed, np.float64)
        else:
            vmax = None
        return mass, size, accel, vmax


HD_PARAMS = WorldParams(agent_size=0.03, world_length=100)
BASIC_PARAMS = WorldParams(agent_size=0.1, world_length=50)


def _two_prod(a, b):
    """Error-free product (Dekker/Veltkamp): a*b == p + e exactly."""
    p = a * b
    c = 134217729.0
    ah = a * c
    ah = ah - (ah - a)
    al = a - ah
    bh = b * c
    bh = bh - (bh - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def norm2(x0, x1):
    """``np.linalg.norm([x0, x1])`` as the reference evaluates it (core.py:305,
    formation_hd_env.py:69,120, basic_formation_env.py:46,90): for a 1-D vector numpy computes
    ``sqrt(dot(x, x))`` and the BLAS ddot of the numpy build the golden files were generated with
    contracts the 2-term sum into ``fma(x1, x1, x0*x0)``.  Emulated here with error-free
    transformations (verified equal to np.linalg.norm on 6e5 random vectors, 0 mismatches); the
    difference from the plain ``sqrt(x0*x0 + x1*x1)`` is at most 1 ulp, which only matters for
    bit-level agreement of long stiff-contact trajectories."""
    with np.errstate(all='ignore'):
        a = x0 * x0
        p, e = _two_prod(x1, x1)
        s = a + p
        bb = s - a
        t = (a - (s - bb)) + (p - bb)
        r = np.sqrt(s + (t + e))
    # inf/nan inputs: fall back to the plain formula (error terms are nan there)
    plain = np.sqrt(x0 * x0 + x1 * x1)
    return np.where(np.isfinite(r), r, plain)


# ------------------------------------------------------------------------------------------
# World.step  (core.py:206-225)
# ------------------------------------------------------------------------------------------
def wall_force(p, size, wall, prm):
    """get_wall_collision_force (core.py:325-362) for a batch of positions p [E,2]."""
    orient, axis_pos, e0, e1, width, _hard = wall
    prll, perp = (0, 1) if orient == 'H' else (1, 0)
    x = p[:, prll]
    beyond = (x < e0 - size) | (x > e1 + size)
    partial = ((x < e0) | (x > e1)) & ~beyond
    past = np.where(x < e0, x - e0, x - e1)
    past = np.where(partial, past, 0.0)
    theta = np.arcsin(past / size)
    dist_min = np.where(partial, np.cos(theta) * size + 0.5 * width, size + 0.5 * width)
    delta = p[:, perp] - axis_pos
    dist = np.abs(delta)
    k = prm.contact_margin
    with np.errstate(all='ignore'):
        pen = np.logaddexp(0, -(dist - dist_min) / k) * k
        fmag = prm.contact_force * delta / dist * pen
    f = np.zeros_like(p)
    f[:, perp] = np.cos(theta) * fmag
    f[:, prll] = np.sin(theta) * np.abs(fmag)
    f[beyond] = 0.0
    return f


def world_step(pos, vel, act, prm=HD_PARAMS, noise=None):
    """One ``World.step`` after ``_set_action``.  Returns new (pos, vel); inputs untouched.

    environment.py:216-221 (u = a * sensitivity), core.py:228-237 (action force),
    core.py:240-262 + 289-322 (pairwise contact force from the OLD positions),
    core.py:264-277 (damping, F/m*dt, max_speed clamp, p += v*dt)."""
    pos = np.asarray(pos, np.float64)
    vel = np.asarray(vel, np.float64)
    act = np.asarray(act, np.float64)
    E, N, _ = pos.shape
    mass, size, accel, vmax = prm.per_agent(N)
    sens = np.full(N, prm.sensitivity) if accel is None else accel        # environment.py:218-220
    gain = mass if accel is None else mass * accel                         # core.py:235-236
    u = act * sens[None, :, None]
    F = gain[None, :, None] * u + (0.0 if noise is None else noise)        # core.py:232-236
    F = np.array(F, np.float64)
    if prm.collide:
        k = prm.contact_margin
        with np.errstate(all='ignore'):      # coincident agents -> 0/0 -> NaN, as the reference
            for a in range(N):
                for b in range(a + 1, N):
                    delta = pos[:, a] - pos[:, b]                          # core.py:304
                    dist = norm2(delta[:, 0], delta[:, 1])                      # core.py:305
                    dist_min = size[a] + size[b]                           # core.py:307
                    pen = np.logaddexp(0, -(dist - dist_min) / k) * k      # core.py:310
                    force = prm.contact_force * delta / dist[:, None] * pen[:, None]  # :312
                    ratio = mass[b] / mass[a]                              # core.py:316
                    F[:, a] = ratio * force + F[:, a]                      # core.py:250,317
                    F[:, b] = -(1 / ratio) * force + F[:, b]               # core.py:254,318
    for w in prm.walls:                                                    # core.py:255-261
        for a in range(N):
            F[:, a] = F[:, a] + wall_force(pos[:, a], size[a], w, prm)
    v = vel * (1 - prm.damping)                                            # core.py:268
    v = v + (F / mass[None, :, None]) * prm.dt                             # core.py:270
    if vmax is not None:                                                   # core.py:271-276
        sp = np.sqrt(np.square(v[..., 0]) + np.square(v[..., 1]))
        over = sp > vmax[None, :]
        with np.errstate(all='ignore'):
            clamped = v / sp[..., None] * vmax[None, :, None]
        v = np.where(over[..., None], clamped, v)
    p = pos + v * prm.dt                                                   # core.py:277
    return p, v


# ------------------------------------------------------------------------------------------
# formation_hd_env scenario hooks
# ------------------------------------------------------------------------------------------
def hd_observation(pos, vel, ideal_shape, ideal_vel):
    """formation_hd_env.py:52-59: [v_i, (p_j - p_i) j!=i ascending, comm zeros 2(N-1),
    ideal_shape.flatten(), ideal_vel] -> [E,N,6N]."""
    E, N, _ = pos.shape
    obs = np.zeros((E, N, 6 * N), np.float64)
    flat_shape = np.asarray(ideal_shape, np.float64).reshape(E, 2 * N)
    for i in range(N):
        obs[:, i, 0:2] = vel[:, i]
        others = [j for j in range(N) if j != i]
        rel = pos[:, others] - pos[:, i:i + 1]
        obs[:, i, 2:2 * N] = rel.reshape(E, 2 * (N - 1))
        obs[:, i, 4 * N - 2:6 * N - 2] = flat_shape
        obs[:, i, 6 * N - 2:6 * N] = ideal_vel
    return obs


def hd_landmark_shift(pos, landmarks, times=1):
    """Side effect of every hd ``observation`` call (formation_hd_env.py:40-44): landmarks are
    re-centred on the agents' centroid.  Called N times per step/reset by the env facade."""
    lm = np.array(landmarks, np.float64)
    for _ in range(times):
        delta = np.mean(pos, 1) - np.mean(lm, 1)
        lm = lm + delta[:, None, :]
    return lm


def directed_hausdorff_sq(u, v):
    """max_i min_j |u_i - v_j|^2 for batches u [E,N,2], v [E,M,2] (scipy's result squared)."""
    dx = u[:, :, None, 0] - v[:, None, :, 0]
    dy = u[:, :, None, 1] - v[:, None, :, 1]
    d2 = dx * dx + dy * dy
    return d2.min(2).max(1)


def hd_reward(pos, vel, ideal_shape, ideal_vel, prm=HD_PARAMS):
    """formation_hd_env.py:61-75,119-121 -> individual rewards [E,N] (post-step state)."""
    E, N, _ = pos.shape
    _, size, _, _ = prm.per_agent(N)
    C = pos - np.mean(pos, 1)[:, None, :]                                   # :64-65
    S = np.asarray(ideal_shape, np.float64)
    form = -np.sqrt(np.maximum(directed_hausdorff_sq(C, S), directed_hausdorff_sq(S, C)))  # :66
    mean_vel = np.mean(vel, 1)                                              # :68
    dv = np.asarray(ideal_vel, np.float64) - mean_vel
    velr = norm2(dv[:, 0], dv[:, 1])                                        # :69
    base = form - velr
    rew = np.repeat(base[:, None], N, 1)
    if prm.collide:                                                         # :71-74
        for i in range(N):
            r = rew[:, i].copy()
            for j in range(N):
                if j == i:
                    continue
                d = pos[:, j] - pos[:, i]
                dist = norm2(d[:, 0], d[:, 1])
                r = np.where(dist < (size[i] + size[j]) / 2, r - 1, r)      # :119-121
            rew[:, i] = r
    return rew


# ------------------------------------------------------------------------------------------
# basic_formation_env scenario hooks
# ------------------------------------------------------------------------------------------
def basic_observation(pos, vel, landmarks):
    """basic_formation_env.py:29-41: [v_i, p_i, (l_k - p_i) k, (p_j - p_i) j!=i, comm zeros]."""
    E, N, _ = pos.shape
    L = landmarks.shape[1]
    D = 4 + 2 * L + 4 * (N - 1)
    obs = np.zeros((E, N, D), np.float64)
    for i in range(N):
        obs[:, i, 0:2] = vel[:, i]
        obs[:, i, 2:4] = pos[:, i]
        obs[:, i, 4:4 + 2 * L] = (landmarks - pos[:, i:i + 1]).reshape(E, 2 * L)
        others = [j for j in range(N) if j != i]
        obs[:, i, 4 + 2 * L:4 + 2 * L + 2 * (N - 1)] = \
            (pos[:, others] - pos[:, i:i + 1]).reshape(E, 2 * (N - 1))
    return obs


def basic_reward(pos, landmarks, prm=BASIC_PARAMS):
    """basic_formation_env.py:43-52,89-91: -sum_l min_a |p_a - l| - #{a (incl. self): |p_a-p_i| < s_a+s_i}."""
    E, N, _ = pos.shape
    L = landmarks.shape[1]
    _, size, _, _ = prm.per_agent(N)
    base = np.zeros(E, np.float64)
    for k in range(L):
        d = pos - landmarks[:, k:k + 1]
        dist = norm2(d[..., 0], d[..., 1])
        base = base - dist.min(1)
    rew = np.repeat(base[:, None], N, 1)
    if prm.collide:
        for i in range(N):
            r = rew[:, i].copy()
            for j in range(N):                      # includes j == i (dist 0 < 2s): Q7
                d = pos[:, j] - pos[:, i]
                dist = norm2(d[:, 0], d[:, 1])
                r = np.where(dist < (size[i] + size[j]), r - 1, r)
            rew[:, i] = r
    return rew


# ------------------------------------------------------------------------------------------
# formation_hd_partial_env / formation_hd_partial_range_env scenario hooks
# ------------------------------------------------------------------------------------------
PARTIAL_PARAMS = WorldParams(agent_size=0.04, world_length=25)   # formation_hd_partial_env.py:15,29


def partial_observation(pos, vel, landmarks, num_obs):
    """formation_hd_partial_env.py:41-66: [v_i, landmark positions (absolute), p_j - p_i for
    j = i+1 .. i+num_obs (cyclic), comm of the others (zeros)]."""
    E, N, _ = pos.shape
    L = landmarks.shape[1]
    D = 2 + 2 * L + 2 * num_obs + 2 * (N - 1)
    obs = np.zeros((E, N, D), np.float64)
    for i in range(N):
        obs[:, i, 0:2] = vel[:, i]
        obs[:, i, 2:2 + 2 * L] = landmarks.reshape(E, 2 * L)
        idx = [j % N for j in range(i + 1, i + 1 + num_obs)]                 # :53
        obs[:, i, 2 + 2 * L:2 + 2 * L + 2 * num_obs] = (pos[:, idx] - pos[:, i:i + 1]).reshape(E, 2 * num_obs)
    return obs


def range_observation(pos, vel, landmarks, obs_range):
    """formation_hd_partial_range_env.py:41-54: other_pos of ALL others clipped to +-obs_range."""
    E, N, _ = pos.shape
    L = landmarks.shape[1]
    D = 2 + 2 * L + 4 * (N - 1)
    obs = np.zeros((E, N, D), np.float64)
    for i in range(N):
        obs[:, i, 0:2] = vel[:, i]
        obs[:, i, 2:2 + 2 * L] = landmarks.reshape(E, 2 * L)
        others = [j for j in range(N) if j != i]
        rel = np.clip(pos[:, others] - pos[:, i:i + 1], -obs_range, obs_range)   # :53
        obs[:, i, 2 + 2 * L:2 + 2 * L + 2 * (N - 1)] = rel.reshape(E, 2 * (N - 1))
    return obs


def partial_reward(pos, landmarks, prm=PARTIAL_PARAMS):
    """formation_hd_partial_env.py:68-87,123-125 (same in the range variant :56-75,111-113):
    -max(dH(u, v), dH(v, u)) with u = agents - mean, v = landmarks - mean; -1 per other agent closer
    than s_a + s_i."""
    E, N, _ = pos.shape
    _, size, _, _ = prm.per_agent(N)
    u = pos - np.mean(pos, 1)[:, None, :]
    v = landmarks - np.mean(landmarks, 1)[:, None, :]
    base = -np.sqrt(np.maximum(directed_hausdorff_sq(u, v), directed_hausdorff_sq(v, u)))
    rew = np.repeat(base[:, None], N, 1)
    if prm.collide:
        for i in range(N):
            r = rew[:, i].copy()
            for j in range(N):
                if j == i:
                    continue
                d = pos[:, j] - pos[:, i]
                dist = norm2(d[:, 0], d[:, 1])
                r = np.where(dist < (size[i] + size[j]), r - 1, r)
            rew[:, i] = r
    return rew


def partial_env_step(pos, vel, act, landmarks, step, num_obs=None, obs_range=None, prm=PARTIAL_PARAMS,
                     noise=None):
    """env.step for formation_hd_partial_env (num_obs given) or formation_hd_partial_range_env (obs_range)."""
    p, v = world_step(pos, vel, act, prm, noise)
    obs = partial_observation(p, v, landmarks, num_obs) if num_obs is not None \
        else range_observation(p, v, landmarks, obs_range)
    indiv = partial_reward(p, landmarks, prm)
    new_step = np.asarray(step) + 1
    return dict(pos=p, vel=v, obs=obs, indiv=indiv, reward=shared_reward(indiv),
                done=new_step >= prm.world_length, step=new_step)


# ------------------------------------------------------------------------------------------
# formation_hd_obs_env: movable colliding obstacle landmarks
# ------------------------------------------------------------------------------------------
OBSTACLE_PARAMS = WorldParams(agent_size=0.1, world_length=50)    # formation_hd_obs_env.py:14,29
OBSTACLE_SIZE = 0.15                                               # formation_hd_obs_env.py:44
OBSTACLE_FLOOR = -2.2                                              # formation_hd_obs_env.py:86


def obstacle_world_step(pos, vel, act, obst, obst_vel, prm=OBSTACLE_PARAMS, obstacle_size=OBSTACLE_SIZE,
                        obstacle_mass=1.0, noise=None):
    """``World.step`` with movable colliding landmarks (core.py:206-277): the entity list is agents,
    goal landmarks, obstacles (core.py:143-144, formation_hd_obs_env.py:31-44); goal landmarks do not collide
    (core.py:292) and drop out.  Pairs a<b over [agents, obstacles]: agent-agent, agent-obstacle,
    obstacle-obstacle, all "both movable" (core.py:314-318).  Obstacles get no action force, are damped and
    integrated like agents (core.py:264-277), never speed-clamped (max_speed None).
    Returns new (pos, vel, obst, obst_vel)."""
    pos = np.asarray(pos, np.float64); vel = np.asarray(vel, np.float64); act = np.asarray(act, np.float64)
    obst = np.asarray(obst, np.float64); obst_vel = np.asarray(obst_vel, np.float64)
    E, N, _ = pos.shape
    O = obst.shape[1]
    mass, size, accel, vmax = prm.per_agent(N)
    sens = np.full(N, prm.sensitivity) if accel is None else accel
    gain = mass if accel is None else mass * accel
    F = np.array(gain[None, :, None] * (act * sens[None, :, None]) + (0.0 if noise is None else noise), np.float64)
    P = np.concatenate([pos, obst], 1)
    Fall = np.concatenate([F, np.zeros((E, O, 2))], 1)            # p_force[b] = 0.0 before the first add (:252-253)
    esize = np.concatenate([size, np.full(O, obstacle_size)])
    emass = np.concatenate([mass, np.full(O, obstacle_mass)])
    agents_collide = [prm.collide] * N + [True] * O
    k = prm.contact_margin
    with np.errstate(all='ignore'):
        for a in range(N + O):
            for b in range(a + 1, N + O):
                if not (agents_collide[a] and agents_collide[b]):
                    continue
                delta = P[:, a] - P[:, b]
                dist = norm2(delta[:, 0], delta[:, 1])
                dist_min = esize[a] + esize[b]
                pen = np.logaddexp(0, -(dist - dist_min) / k) * k
                force = prm.contact_force * delta / dist[:, None] * pen[:, None]
                ratio = emass[b] / emass[a]
                Fall[:, a] = ratio * force + Fall[:, a]
                Fall[:, b] = -(1 / ratio) * force + Fall[:, b]
    for w in prm.walls:                                            # every movable entity (core.py:255-261)
        for a in range(N + O):
            Fall[:, a] = Fall[:, a] + wall_force(P[:, a], esize[a], w, prm)
    V = np.concatenate([vel, obst_vel], 1) * (1 - prm.damping)
    V = V + (Fall / emass[None, :, None]) * prm.dt
    if vmax is not None:
        va = V[:, :N]
        sp = np.sqrt(np.square(va[..., 0]) + np.square(va[..., 1]))
        with np.errstate(all='ignore'):
            clamped = va / sp[..., None] * vmax[None, :, None]
        V[:, :N] = np.where((sp > vmax[None, :])[..., None], clamped, va)
    P = P + V * prm.dt
    return P[:, :N], V[:, :N], P[:, N:], V[:, N:]


def obstacle_observation(pos, vel, goals, obst):
    """formation_hd_obs_env.py:53-68: [v_i, goal landmark positions (absolute), obstacle positions - p_i,
    p_j - p_i (j != i), comm of the others (zeros)]."""
    E, N, _ = pos.shape
    L, O = goals.shape[1], obst.shape[1]
    D = 2 + 2 * L + 2 * O + 4 * (N - 1)
    obs = np.zeros((E, N, D), np.float64)
    for i in range(N):
        obs[:, i, 0:2] = vel[:, i]
        obs[:, i, 2:2 + 2 * L] = goals.reshape(E, 2 * L)
        obs[:, i, 2 + 2 * L:2 + 2 * L + 2 * O] = (obst - pos[:, i:i + 1]).reshape(E, 2 * O)
        others = [j for j in range(N) if j != i]
        b = 2 + 2 * L + 2 * O
        obs[:, i, b:b + 2 * (N - 1)] = (pos[:, others] - pos[:, i:i + 1]).reshape(E, 2 * (N - 1))
    return obs


def obstacle_reward(pos, goals, obst, prm=OBSTACLE_PARAMS, obstacle_size=OBSTACLE_SIZE):
    """formation_hd_obs_env.py:70-99,147-149: -max(dH(u, v), dH(v, u)) with u = agents - mean, v = goal
    landmarks - mean; -2 per other agent closer than s_a + s_i; -2 per obstacle closer than s_o + s_i."""
    E, N, _ = pos.shape
    _, size, _, _ = prm.per_agent(N)
    u = pos - np.mean(pos, 1)[:, None, :]
    v = goals - np.mean(goals, 1)[:, None, :]
    base = -np.sqrt(np.maximum(directed_hausdorff_sq(u, v), directed_hausdorff_sq(v, u)))
    rew = np.repeat(base[:, None], N, 1)
    if prm.collide:
        for i in range(N):
            r = rew[:, i].copy()
            for j in range(N):
                if j == i:
                    continue
                d = pos[:, j] - pos[:, i]
                r = np.where(norm2(d[:, 0], d[:, 1]) < (size[i] + size[j]), r - 2, r)
            for kk in range(obst.shape[1]):
                d = obst[:, kk] - pos[:, i]
                r = np.where(norm2(d[:, 0], d[:, 1]) < (obstacle_size + size[i]), r - 2, r)
            rew[:, i] = r
    return rew


def obstacle_velocity_rule(obst):
    """Side effect of every ``reward`` call (formation_hd_obs_env.py:85-88): an obstacle above the floor falls at
    (0, -1), below it stops."""
    v = np.zeros_like(obst)
    v[..., 1] = np.where(obst[..., 1] > OBSTACLE_FLOOR, -1.0, 0.0)
    return v


def obstacle_env_step(pos, vel, act, goals, obst, obst_vel, step, prm=OBSTACLE_PARAMS,
                      obstacle_size=OBSTACLE_SIZE, noise=None):
    """env.step for formation_hd_obs_env (environment.py:113-142 over the hooks above)."""
    p, v, o, ov = obstacle_world_step(pos, vel, act, obst, obst_vel, prm, obstacle_size, 1.0, noise)
    obs = obstacle_observation(p, v, np.asarray(goals, np.float64), o)
    indiv = obstacle_reward(p, np.asarray(goals, np.float64), o, prm, obstacle_size)
    new_step = np.asarray(step) + 1
    return dict(pos=p, vel=v, obst=o, obst_vel=obstacle_velocity_rule(o), obst_vel_integrated=ov, obs=obs,
                indiv=indiv, reward=shared_reward(indiv), done=new_step >= prm.world_length, step=new_step)


def obstacle_reset_from_uniform(u_pos, u_goals, u_obst):
    """formation_hd_obs_env.py:101-120 from U(0,1) draws ``u_obst`` [E,O,2] and U(-1,1) draws for agents and
    goal landmarks: obstacle k starts at x ~ U(step[k], step[k+1]), step = linspace(-1.8, 1.8, O+1),
    y ~ U(2.0, 2.5), velocity (0, -1)."""
    u_obst = np.asarray(u_obst, np.float64)
    O = u_obst.shape[1]
    st = np.linspace(-1.8, 1.8, O + 1)
    obst = np.empty_like(u_obst)
    obst[..., 0] = st[:-1][None, :] + (st[1:] - st[:-1])[None, :] * u_obst[..., 0]
    obst[..., 1] = 2.0 + 0.5 * u_obst[..., 1]
    ov = np.zeros_like(obst)
    ov[..., 1] = -1.0
    pos = np.array(u_pos, np.float64)
    return pos, np.zeros_like(pos), np.array(u_goals, np.float64), obst, ov


# ------------------------------------------------------------------------------------------
# MultiAgentEnv.step  (environment.py:113-142)
# ------------------------------------------------------------------------------------------
def shared_reward(indiv):
    """environment.py:136-138: reward = np.sum(reward_n) with reward_n a list of N [r_i]."""
    return np.array([np.sum(row.reshape(-1, 1)) for row in indiv], np.float64)


def hd_env_step(pos, vel, act, ideal_shape, ideal_vel, step, prm=HD_PARAMS, landmarks=None,
                noise=None):
    """Full ``env.step`` for formation_hd_env.  ``step`` [E] is current_step BEFORE the call."""
    p, v = world_step(pos, vel, act, prm, noise)
    obs = hd_observation(p, v, ideal_shape, ideal_vel)
    indiv = hd_reward(p, v, ideal_shape, ideal_vel, prm)
    new_step = np.asarray(step) + 1                                         # environment.py:114
    done = new_step >= prm.world_length                                     # :172-177
    out = dict(pos=p, vel=v, obs=obs, indiv=indiv, reward=shared_reward(indiv),
               done=done, step=new_step)
    if landmarks is not None:
        out['landmarks'] = hd_landmark_shift(p, landmarks, times=p.shape[1])
    return out


def basic_env_step(pos, vel, act, landmarks, step, prm=BASIC_PARAMS, noise=None):
    p, v = world_step(pos, vel, act, prm, noise)
    obs = basic_observation(p, v, landmarks)
    indiv = basic_reward(p, landmarks, prm)
    new_step = np.asarray(step) + 1
    done = new_step >= prm.world_length
    return dict(pos=p, vel=v, obs=obs, indiv=indiv, reward=shared_reward(indiv),
                done=done, step=new_step)


# ------------------------------------------------------------------------------------------
# reset (formation_hd_env.py:77-95, basic_formation_env.py:54-65) from given uniforms
# ------------------------------------------------------------------------------------------
def hd_reset_from_uniform(u_pos, u_lm, u_vel):
    """u_* are U(-1,1) draws in the reference's order (agents, landmarks, ideal_vel).
    Returns pos, vel, landmarks, ideal_shape (centred, :93), ideal_vel."""
    pos = np.array(u_pos, np.float64)
    lm = np.array(u_lm, np.float64)
    shape = lm - np.mean(lm, 1)[:, None, :]
    return pos, np.zeros_like(pos), lm, shape, np.array(u_vel, np.float64)
