"""Device backend of the single-env drop-in facade (E = 1).

Gathers the host-side ``World`` records (formation_gym/core.py data model) into ``[1,N,2]`` device
tensors, calls the C ABI (include/formation_gym_b200.h) and scatters the results back, so that
code written against the reference API (``env.world.agents[i].state.p_pos`` ...) keeps working
while all arithmetic of the step path runs in the sm_100a kernels.  fp64 kernels by default
(the reference computes in float64).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat


def _uniform(values, what, allow_mixed=False):
    vals = list(values)
    same = all((v == vals[0]) or (v is None and vals[0] is None) for v in vals)
    if same:
        return vals[0], None
    if not allow_mixed:
        raise NotImplementedError("per-agent %s is not supported by the accelerated path" % what)
    return None, vals


class FacadeBackend(object):
    def __init__(self, world):
        self.lib = nat.load()
        if not torch.cuda.is_available():
            raise nat.NativeError("no CUDA device: formation_gym has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.N = len(world.agents)
        self.L = len(world.landmarks)
        self.np_dtype = np.dtype(world.dtype)
        self.dtype = torch.float64 if self.np_dtype == np.float64 else torch.float32
        self.sfx = "_f64" if self.dtype == torch.float64 else ""
        N, L = self.N, self.L
        kw = dict(device=self.device, dtype=self.dtype)
        self.pos = torch.zeros(1, N, 2, **kw)
        self.vel = torch.zeros(1, N, 2, **kw)
        self.act = torch.zeros(1, N, 4, **kw)
        self.comm = torch.zeros(1, N, 2, **kw)
        self.shape = torch.zeros(1, N, 2, **kw)
        self.ivel = torch.zeros(1, 2, **kw)
        self.lm = torch.zeros(1, max(L, 1), 2, **kw)
        self.lmv = torch.zeros(1, max(L, 1), 2, **kw)          # landmark.state.p_vel (obstacle scenario)
        self._n_obst = 0
        self.step = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.obs = None
        self.reward = torch.zeros(1, N, 1, **kw)
        self.indiv = torch.zeros(1, N, **kw)
        self.done = torch.zeros(1, N, dtype=torch.uint8, device=self.device)
        self._keep = []
        self._cache_key = None
        self._cache_val = None
        self.launches = 0

    def matches(self, world):
        return (self.N == len(world.agents) and self.L == len(world.landmarks)
                and self.np_dtype == np.dtype(world.dtype))

    # ------------------------------------------------------------------ host <-> device
    def _up(self, dst, arr):
        a = np.ascontiguousarray(arr, dtype=self.np_dtype).reshape(tuple(dst.shape))
        dst.copy_(torch.from_numpy(a))

    def _params(self, world, scenario_kind, prescaled, scenario=None):
        agents = world.agents
        if not all(a.movable for a in agents):
            raise NotImplementedError("immovable agents are not supported by the accelerated path")
        obstacles = [l for l in world.landmarks if l.movable or l.collide]
        n_obst = len(obstacles)
        if n_obst:
            # formation_hd_obs_env: trailing landmarks that are movable AND collide, uniform size / mass
            tail = world.landmarks[len(world.landmarks) - n_obst:]
            ok = (scenario_kind == nat.FG_SCENARIO_HD_OBSTACLE and n_obst < len(world.landmarks)
                  and all(a is b for a, b in zip(tail, obstacles))
                  and all(l.movable and l.collide and l.max_speed is None for l in obstacles)
                  and len({(float(l.size), float(l.mass)) for l in obstacles}) == 1)
            if not ok:
                raise NotImplementedError(
                    "movable / colliding landmarks are supported as the trailing, uniform obstacles of the "
                    "formation_hd_obs_env scenario hooks only (fg_step_fused / fg_obs_reward)")
        elif scenario_kind == nat.FG_SCENARIO_HD_OBSTACLE and len(world.landmarks) < 1:
            raise NotImplementedError("formation_hd_obs_env needs at least one goal landmark")
        collide, _ = _uniform([bool(a.collide) for a in agents], "collide")
        silent, _ = _uniform([bool(a.silent) for a in agents], "silent")
        u_noise, _ = _uniform([a.u_noise or None for a in agents], "u_noise")
        c_noise, _ = _uniform([a.c_noise or None for a in agents], "c_noise")
        if not silent and world.dim_c != 2:
            raise NotImplementedError("non-silent agents need dim_c == 2")
        mass, mass_arr = _uniform([float(a.mass) for a in agents], "mass", True)
        size, size_arr = _uniform([float(a.size) for a in agents], "size", True)
        accel, accel_arr = _uniform([a.accel for a in agents], "accel", True)
        vmax, vmax_arr = _uniform([a.max_speed for a in agents], "max_speed", True)
        walls = [(w.orient, float(w.axis_pos), float(w.endpoints[0]), float(w.endpoints[1]),
                  float(w.width), bool(w.hard)) for w in world.walls]
        if any((not w[5]) for w in walls) and any(getattr(a, "ghost", False) for a in agents):
            raise NotImplementedError("ghost agents with soft walls are not supported")
        p = nat.make_params(
            dt=world.dt, damping=world.damping, contact_force=world.contact_force,
            contact_margin=world.contact_margin, sensitivity=5.0,
            agent_size=size if size_arr is None else size_arr[0],
            mass=mass if mass_arr is None else mass_arr[0],
            accel=accel if accel_arr is None else None,
            max_speed=vmax if vmax_arr is None else None,
            u_noise=u_noise, c_noise=c_noise, collide=collide, silent=silent,
            world_length=world.world_length, walls=walls, action_prescaled=prescaled,
            num_obs=getattr(scenario, "num_obs", 0) or 0, obs_range=getattr(scenario, "obs_range", 0.0) or 0.0,
            num_obstacles=n_obst, obstacle_size=float(obstacles[0].size) if n_obst else 0.15,
            obstacle_mass=float(obstacles[0].mass) if n_obst else 1.0)
        self._n_obst = n_obst
        self._keep = []
        for field, arr in (("agent_mass", mass_arr), ("agent_size_arr", size_arr),
                           ("agent_accel", accel_arr), ("agent_max_speed", vmax_arr)):
            if arr is not None:
                t = torch.tensor([(-1.0 if x is None else float(x)) for x in arr],
                                 device=self.device, dtype=self.dtype)
                self._keep.append(t)
                setattr(p, field, t.data_ptr())
        return p, silent

    def _gather_state(self, world):
        P = np.stack([np.asarray(a.state.p_pos, np.float64) for a in world.agents])
        V = np.stack([np.asarray(a.state.p_vel, np.float64) for a in world.agents])
        return P, V

    def _scatter_state(self, world, with_comm=True):
        P = self.pos[0].cpu().numpy().astype(np.float64)
        V = self.vel[0].cpu().numpy().astype(np.float64)
        Cm = self.comm[0].cpu().numpy().astype(np.float64)
        for i, a in enumerate(world.agents):
            a.state.p_pos = P[i].copy()
            a.state.p_vel = V[i].copy()
            if with_comm:
                a.state.c = Cm[i, :world.dim_c].copy() if world.dim_c <= 2 else np.zeros(world.dim_c)

    def _buffers(self, scenario_kind, with_obs):
        b = nat.fg_buffers()
        b.pos, b.vel, b.act, b.comm = nat.ptr(self.pos), nat.ptr(self.vel), nat.ptr(self.act), nat.ptr(self.comm)
        b.ideal_shape, b.ideal_vel = nat.ptr(self.shape), nat.ptr(self.ivel)
        b.landmarks = nat.ptr(self.lm) if self.L > 0 else None
        b.landmark_vel = nat.ptr(self.lmv) if (self.L > 0 and scenario_kind == nat.FG_SCENARIO_HD_OBSTACLE) else None
        b.step = nat.ptr(self.step)
        if with_obs:
            from .batched import obs_dim, SCENARIOS
            name = [k for k, v in SCENARIOS.items() if v == scenario_kind][0]
            D = obs_dim(name, self.N, self.L, getattr(self, "_num_obs", 3))
            if self.obs is None or self.obs.shape[2] != D:
                self.obs = torch.zeros(1, self.N, D, device=self.device, dtype=self.dtype)
            b.obs = nat.ptr(self.obs)
        b.reward, b.indiv, b.done = nat.ptr(self.reward), nat.ptr(self.indiv), nat.ptr(self.done)
        return b

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _upload_actions(self, U, silent, Cact=None):
        if silent:
            act = self.act.view(-1)[: self.N * 2].view(1, self.N, 2)
            self._up(act, U)
        else:
            self._up(self.act, np.concatenate([U, Cact], axis=1))

    # ------------------------------------------------------------------ World.step
    def world_step(self, world):
        """core.py:206-225 on the GPU; ``agent.action.u`` is already scaled (action_prescaled)."""
        p, silent = self._params(world, nat.FG_SCENARIO_BASIC, True)
        P, V = self._gather_state(world)
        U = np.stack([np.asarray(a.action.u, np.float64) for a in world.agents])
        Cact = None if silent else np.stack([np.asarray(a.action.c, np.float64) for a in world.agents])
        self._up(self.pos, P)
        self._up(self.vel, V)
        self._upload_actions(U, silent, Cact)
        b = self._buffers(nat.FG_SCENARIO_BASIC, False)
        fn = getattr(self.lib, "fg_world_step" + self.sfx)
        with torch.cuda.device(self.device):
            nat.check(fn(C.byref(p), C.byref(b), 1, self.N, int(world.seed) & (2 ** 64 - 1),
                         int(world.world_step) & 0xFFFFFFFF, 0, self._stream()), "fg_world_step")
        self.launches += 1
        self._scatter_state(world)
        self._cache_key = None

    # ------------------------------------------------------------------ scenario hooks
    def _upload_scenario(self, world, scenario, kind):
        if kind == nat.FG_SCENARIO_HD:
            self._up(self.shape, np.asarray(scenario.ideal_shape, np.float64))
            self._up(self.ivel, np.asarray(scenario.ideal_vel, np.float64))
        if self.L > 0:
            self._up(self.lm, np.stack([np.asarray(l.state.p_pos, np.float64) for l in world.landmarks]))
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self._up(self.lmv, np.stack([np.zeros(2) if l.state.p_vel is None else np.asarray(l.state.p_vel, np.float64)
                                         for l in world.landmarks]))

    def _scatter_obstacles(self, world, with_pos):
        """Obstacle state back to the host records: positions after a fused step, velocities as the reward
        hook's rule leaves them (formation_hd_obs_env.py:85-88)."""
        n = self._n_obst
        if not n:
            return
        lm = self.lm[0].cpu().numpy().astype(np.float64)
        lv = self.lmv[0].cpu().numpy().astype(np.float64)
        for k in range(self.L - n, self.L):
            if with_pos:
                world.landmarks[k].state.p_pos = lm[k].copy()
            world.landmarks[k].state.p_vel = lv[k].copy()

    def scenario_eval(self, world, scenario, kind):
        """observation + reward of every agent from the CURRENT host state (one launch, cached on
        the state's bytes so the 3N hook calls of one env.step share it)."""
        P, V = self._gather_state(world)
        Cm = np.stack([np.asarray(a.state.c, np.float64) if a.state.c is not None
                       else np.zeros(2) for a in world.agents]) if world.dim_c == 2 else np.zeros((self.N, 2))
        lmh = np.stack([np.asarray(l.state.p_pos, np.float64) for l in world.landmarks]) if self.L else np.zeros((0, 2))
        key = (P.tobytes(), V.tobytes(), Cm.tobytes(), lmh.tobytes(),
               np.asarray(getattr(scenario, "ideal_shape", 0.0), np.float64).tobytes(),
               np.asarray(getattr(scenario, "ideal_vel", 0.0), np.float64).tobytes())
        if key == self._cache_key:
            if kind == nat.FG_SCENARIO_HD_OBSTACLE:
                self._scatter_obstacles(world, with_pos=False)
            return self._cache_val
        p, _ = self._params(world, kind, False, scenario)
        self._num_obs = int(getattr(scenario, "num_obs", 3) or 0)
        self._up(self.pos, P)
        self._up(self.vel, V)
        self._up(self.comm, Cm)
        self._upload_scenario(world, scenario, kind)
        b = self._buffers(kind, True)
        fn = getattr(self.lib, "fg_obs_reward" + self.sfx)
        with torch.cuda.device(self.device):
            nat.check(fn(C.byref(p), C.byref(b), kind, 1, self.N, self.L, self._stream()), "fg_obs_reward")
        self.launches += 1
        val = dict(obs=self.obs[0].cpu().numpy().astype(np.float64),
                   indiv=self.indiv[0].cpu().numpy().astype(np.float64),
                   reward=float(self.reward[0, 0, 0].item()))
        if kind == nat.FG_SCENARIO_HD and self.L > 0:
            # observation's side effect (formation_hd_env.py:40-44): landmarks re-centred
            lm = self.lm[0].cpu().numpy().astype(np.float64)
            for k, l in enumerate(world.landmarks):
                l.state.p_pos = lm[k].copy()
            lmh = lm
            key = key[:3] + (lmh.tobytes(),) + key[4:]
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self._scatter_obstacles(world, with_pos=False)       # the reward hook's velocity rule
        self._cache_key, self._cache_val = key, val
        return val

    # ------------------------------------------------------------------ fused env.step
    def step_fused(self, world, scenario, kind, acts, current_step, acts_c=None):
        """environment.py:113-142 in one launch.  ``current_step`` is the value BEFORE the step."""
        p, silent = self._params(world, kind, False, scenario)
        self._num_obs = int(getattr(scenario, "num_obs", 3) or 0)
        P, V = self._gather_state(world)
        self._up(self.pos, P)
        self._up(self.vel, V)
        self._upload_actions(np.asarray(acts, np.float64), silent, acts_c)
        self._upload_scenario(world, scenario, kind)
        self.step.fill_(int(current_step))
        b = self._buffers(kind, True)
        fn = getattr(self.lib, "fg_step_fused" + self.sfx)
        with torch.cuda.device(self.device):
            nat.check(fn(C.byref(p), C.byref(b), kind, 1, self.N, self.L, 1, 0, 0,
                         int(world.seed) & (2 ** 64 - 1), int(world.world_step) & 0xFFFFFFFF, 0,
                         self._stream()), "fg_step_fused")
        self.launches += 1
        self._scatter_state(world)
        if kind == nat.FG_SCENARIO_HD and self.L > 0:
            lm = self.lm[0].cpu().numpy().astype(np.float64)
            for k, l in enumerate(world.landmarks):
                l.state.p_pos = lm[k].copy()
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self._scatter_obstacles(world, with_pos=True)
        self._cache_key = None
        return dict(obs=self.obs[0].cpu().numpy().astype(np.float64),
                    indiv=self.indiv[0].cpu().numpy().astype(np.float64),
                    reward=float(self.reward[0, 0, 0].item()),
                    done=bool(self.done[0, 0].item()))
