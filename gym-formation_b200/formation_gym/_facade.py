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
from . import core as _core


def _uniform(values, what, allow_mixed=False):
    vals = list(values)
    same = all((v == vals[0]) or (v is None and vals[0] is None) for v in vals)
    if same:
        return vals[0], None
    if not allow_mixed:
        raise NotImplementedError("per-agent %s is not supported by the accelerated path" % what)
    return None, vals


class FacadeBackend(object):
    """One contiguous device ARENA (all buffers of the C ABI, E = 1) with a pinned host mirror: a step costs one
    H2D copy of the inputs, one launch, one D2H copy of the outputs and one stream synchronisation (the first
    version issued ~15 small copies per step: 312 us per basic_formation_env step; the copies dominated)."""

    # (name, elements per unit, unit) in arena order: [in-only | in/out | out-only]; unit 'N' / 'L' / '1' / obs
    _SEGMENTS = (("act", 4, "N"), ("shape", 2, "N"), ("ivel", 2, "1"), ("cpos", 2, "N"),
                 ("pos", 2, "N"), ("vel", 2, "N"), ("comm", 2, "N"), ("lm", 2, "L"), ("lmv", 2, "L"), ("step", 0, "i32"),
                 ("reward", 1, "N"), ("indiv", 1, "N"), ("done", 0, "u8"), ("obs", 0, "obs"))
    _INOUT_FIRST, _OUT_FIRST = "pos", "reward"
    ZERO_COPY_MAX_BYTES = 64 * 1024

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
        self._n_obst = 0
        self._obs_cap = 0
        self._alloc(6 * self.N)
        self._keep = []
        self._pkey = None
        self._pval = None
        self._cache_key = None
        self._cache_val = None
        self._bufcache = {}
        self._cache_pos = None               # World.cache_dists: agent positions at the last calculate_distances()
        self._dist_bufs = None
        self.launches = 0

    def _alloc(self, obs_dim):
        """(Re)build the arena for observation rows of up to ``obs_dim`` scalars."""
        es, N, L = self.np_dtype.itemsize, self.N, max(self.L, 1)
        off, self._off = 0, {}
        for name, per, unit in self._SEGMENTS:
            nbytes = {"N": per * N * es, "L": per * L * es, "1": per * es, "i32": 4, "u8": N,
                      "obs": N * obs_dim * es}[unit]
            self._off[name] = (off, nbytes)
            off = (off + nbytes + 15) & ~15
        self._obs_cap = obs_dim
        self.h = torch.zeros(off, dtype=torch.uint8).pin_memory()
        # Small arenas (a few KB at N = 3 .. 27) are used ZERO-COPY: the kernels read and write the pinned host arena
        # directly through its unified address, so a step is one launch + one stream synchronisation instead of
        # H2D copy + launch + D2H copy + synchronisation (two copy-engine round trips and two API calls less; the
        # step is pure latency at E = 1).  Large arenas (N = 243: 2.8 MB of observations) keep the device copy:
        # streaming them over PCIe from the SMs would be slower than the copy engine.
        self.zero_copy = off <= self.ZERO_COPY_MAX_BYTES
        self.d = self.h if self.zero_copy else torch.zeros(off, dtype=torch.uint8, device=self.device)
        self._base = self.d.data_ptr()
        hn = self.h.numpy()

        def views(name, dtype_t, dtype_n, shape):
            o, nb = self._off[name]
            return (None, hn[o:o + nb].view(dtype_n).reshape(shape))
        T, Tn = self.dtype, self.np_dtype
        self.act, self.h_act = views("act", T, Tn, (1, N, 4))
        self.shape, self.h_shape = views("shape", T, Tn, (1, N, 2))
        self.ivel, self.h_ivel = views("ivel", T, Tn, (1, 2))
        _, self.h_cpos = views("cpos", T, Tn, (1, N, 2))
        self.pos, self.h_pos = views("pos", T, Tn, (1, N, 2))
        self.vel, self.h_vel = views("vel", T, Tn, (1, N, 2))
        self.comm, self.h_comm = views("comm", T, Tn, (1, N, 2))
        self.lm, self.h_lm = views("lm", T, Tn, (1, L, 2))
        self.lmv, self.h_lmv = views("lmv", T, Tn, (1, L, 2))
        self.step, self.h_step = views("step", torch.int32, np.int32, (1,))
        self.reward, self.h_reward = views("reward", T, Tn, (1, N, 1))
        self.indiv, self.h_indiv = views("indiv", T, Tn, (1, N))
        self.done, self.h_done = views("done", torch.uint8, np.uint8, (1, N))
        self.obs, self.h_obs_flat = views("obs", T, Tn, (N * obs_dim,))

    def matches(self, world):
        return (self.N == len(world.agents) and self.L == len(world.landmarks)
                and self.np_dtype == np.dtype(world.dtype))

    # ------------------------------------------------------------------ host <-> device
    def _h2d(self):
        """Everything the kernels read: [act .. step]."""
        if self.zero_copy:
            return
        end = self._off["step"][0] + 16
        self.d[:end].copy_(self.h[:end], non_blocking=True)

    def _d2h(self, last):
        """Everything the kernels wrote, from ``pos`` up to and including segment ``last`` (+ used obs rows)."""
        if not self.zero_copy:
            start = self._off[self._INOUT_FIRST][0]
            o, nb = self._off[last]
            if last == "obs":
                nb = self.N * self._D * self.np_dtype.itemsize
            end = (o + nb + 15) & ~15
            self.h[start:end].copy_(self.d[start:end], non_blocking=True)
        self._sync()

    def _p(self, name):
        """Address the kernels use for arena segment ``name`` (device arena, or the pinned host arena when zero-copy)."""
        return self._base + self._off[name][0]

    def _sync(self):
        torch.cuda.current_stream(self.device).synchronize()

    def _h_obs(self):
        return self.h_obs_flat[: self.N * self._D].reshape(self.N, self._D)

    def _params(self, world, scenario_kind, prescaled, scenario=None):
        """fg_params of the world's constants; rebuilt only when one of them changed."""
        # every attribute assignment on a World / Entity / Wall record bumps core.config_version(); the scenario's
        # own knobs and the (in-place mutable) wall geometry are compared by value
        key = (scenario_kind, bool(prescaled), _core.config_version(), len(world.agents), len(world.landmarks),
               getattr(scenario, "num_obs", None), getattr(scenario, "obs_range", None),
               tuple((w.orient, float(w.axis_pos), float(w.endpoints[0]), float(w.endpoints[1]), float(w.width),
                      bool(w.hard)) for w in world.walls) if world.walls else ())
        if key == self._pkey:
            return self._pval
        self._pval = self._build_params(world, scenario_kind, prescaled, scenario)
        self._pkey = key
        return self._pval

    def _build_params(self, world, scenario_kind, prescaled, scenario=None):
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
                    "movable / colliding landmarks are supported as the trailing, uniform obstacles of a world "
                    "(formation_hd_obs_env layout: goal landmarks first, obstacles last, one size and mass)")
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
            obstacle_mass=float(obstacles[0].mass) if n_obst else 1.0, num_landmarks=len(world.landmarks))
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

    def _stage_state(self, world):
        """Agent positions / velocities into the pinned mirror; returns them (float64) for cache keys."""
        P, V = self._gather_state(world)
        self.h_pos[0] = P
        self.h_vel[0] = V
        return P, V

    def _scatter_state(self, world, with_comm=True):
        P = self.h_pos[0].astype(np.float64)
        V = self.h_vel[0].astype(np.float64)
        Cm = self.h_comm[0].astype(np.float64)
        for i, a in enumerate(world.agents):
            a.state.p_pos = P[i].copy()
            a.state.p_vel = V[i].copy()
            if with_comm:
                a.state.c = Cm[i, :world.dim_c].copy() if world.dim_c <= 2 else np.zeros(world.dim_c)

    def _buffers(self, scenario_kind, with_obs, scenario=None, cached=False):
        """fg_buffers block for one entry point; cached (the arena only moves when it is re-allocated).
        ``cached``: World.cache_dists -- contact forces from the positions of the last calculate_distances()."""
        ck = (scenario_kind, bool(with_obs), int(getattr(scenario, "num_obs", 3) or 0), bool(cached))
        hit = self._bufcache.get(ck)
        if hit is not None and hit[2] == self.d.data_ptr():
            if with_obs:
                self._D = hit[1]
            return hit[0]
        b = self._build_buffers(scenario_kind, with_obs, scenario)
        if cached:
            b.contact_pos = self._p("cpos")
        self._bufcache[ck] = (b, getattr(self, "_D", 0), self.d.data_ptr())
        return b

    def _build_buffers(self, scenario_kind, with_obs, scenario=None):
        b = nat.fg_buffers()
        b.pos, b.vel, b.act, b.comm = self._p("pos"), self._p("vel"), self._p("act"), self._p("comm")
        b.ideal_shape, b.ideal_vel = self._p("shape"), self._p("ivel")
        b.landmarks = self._p("lm") if self.L > 0 else None
        b.landmark_vel = self._p("lmv") if (self.L > 0 and scenario_kind == nat.FG_SCENARIO_HD_OBSTACLE) else None
        b.step = self._p("step")
        if with_obs:
            from .batched import obs_dim, SCENARIOS
            name = [k for k, v in SCENARIOS.items() if v == scenario_kind][0]
            self._D = obs_dim(name, self.N, self.L, int(getattr(scenario, "num_obs", 3) or 0))
            if self._D > self._obs_cap:
                keep = self.h.clone()
                old = dict(self._off)
                self._alloc(self._D)
                for nm, (o, nb) in old.items():                    # staged inputs survive the re-allocation
                    if nm != "obs":
                        o2, _ = self._off[nm]
                        self.h[o2:o2 + nb] = keep[o:o + nb]
                return self._build_buffers(scenario_kind, with_obs, scenario)
            b.obs = self._p("obs")
        b.reward, b.indiv, b.done = self._p("reward"), self._p("indiv"), self._p("done")
        return b

    def _stream(self):
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))

    def _stage_actions(self, U, silent, Cact=None):
        if silent:
            self.h_act.reshape(-1)[: self.N * 2] = np.asarray(U, self.np_dtype).reshape(-1)
        else:
            self.h_act[0] = np.concatenate([U, Cact], axis=1)

    # ------------------------------------------------------------------ World.step
    def world_step(self, world):
        """core.py:206-225 on the GPU; ``agent.action.u`` is already scaled (action_prescaled).  A world whose
        trailing landmarks are movable colliders (formation_hd_obs_env's obstacles) steps them too."""
        has_obst = any(l.movable or l.collide for l in world.landmarks)
        cached = bool(getattr(world, "cache_dists", False))
        if cached:
            if has_obst:
                raise NotImplementedError("World.cache_dists with movable / colliding landmarks is not supported")
            if world.cached_dist_vect is None or self._cache_pos is None:
                # the reference subscripts the still-empty cache here (core.py:299): a world with cache_dists set
                # needs one calculate_distances() call before its first step
                raise TypeError("'NoneType' object is not subscriptable (World.cache_dists: call "
                                "world.calculate_distances() before the first step, core.py:156,299)")
            self.h_cpos[0] = self._cache_pos
        p, silent = self._params(world, nat.FG_SCENARIO_HD_OBSTACLE if has_obst else nat.FG_SCENARIO_BASIC, True)
        self._stage_state(world)
        if has_obst:
            self._stage_scenario(world, None, nat.FG_SCENARIO_HD_OBSTACLE)
        U = np.stack([np.asarray(a.action.u, np.float64) for a in world.agents])
        Cact = None if silent else np.stack([np.asarray(a.action.c, np.float64) for a in world.agents])
        self._stage_actions(U, silent, Cact)
        b = self._buffers(nat.FG_SCENARIO_HD_OBSTACLE if has_obst else nat.FG_SCENARIO_BASIC, False, cached=cached)
        fn = getattr(self.lib, "fg_world_step" + self.sfx)
        with torch.cuda.device(self.device):
            self._h2d()
            nat.check(fn(C.byref(p), C.byref(b), 1, self.N, int(world.seed) & (2 ** 64 - 1),
                         int(world.world_step) & 0xFFFFFFFF, 0, self._stream()), "fg_world_step")
            self._d2h("lmv" if has_obst else "comm")
        self.launches += 1
        self._scatter_state(world)
        if has_obst:
            self._scatter_obstacles(world, with_pos=True)        # integrated velocities: no reward hook ran
        self._cache_key = None

    def calculate_distances(self, world):
        """World.calculate_distances (core.py:156-180) on the device (fg_pair_distances): fills
        ``world.cached_dist_vect [M,M,2]``, ``cached_dist_mag [M,M]``, ``min_dists [M,M]`` and ``cached_collisions [M,M]``
        over world.entities (agents, then landmarks) and remembers the agents' positions of this moment -- the
        positions the NEXT step's contact forces see (core.py:298-301)."""
        ents = world.entities
        M = len(ents)
        P = np.stack([np.asarray(e.state.p_pos, np.float64) for e in ents])
        kw = dict(dtype=self.dtype, device=self.device)
        if self._dist_bufs is None or self._dist_bufs[0].shape[0] != M:
            self._dist_bufs = (torch.empty(M, 2, **kw), torch.empty(M, **kw), torch.empty(M, M, 2, **kw),
                               torch.empty(M, M, **kw), torch.empty(M, M, dtype=torch.uint8, device=self.device),
                               torch.empty(M, M, **kw))
        ent, size, vect, mag, coll, mind = self._dist_bufs
        ent.copy_(torch.as_tensor(P.astype(self.np_dtype)))
        size.copy_(torch.as_tensor(np.array([float(e.size) for e in ents], self.np_dtype)))
        fn = getattr(self.lib, "fg_pair_distances" + self.sfx)
        with torch.cuda.device(self.device):
            nat.check(fn(ent.data_ptr(), size.data_ptr(), 1, M, vect.data_ptr(), mag.data_ptr(), coll.data_ptr(),
                         mind.data_ptr(), self._stream()), "fg_pair_distances")
        self.launches += 1
        world.cached_dist_vect = vect.double().cpu().numpy()
        world.cached_dist_mag = mag.double().cpu().numpy()
        world.min_dists = mind.double().cpu().numpy()
        world.cached_collisions = coll.cpu().numpy().astype(bool)
        self._cache_pos = P[: self.N].copy()

    # ------------------------------------------------------------------ scenario hooks
    def _stage_scenario(self, world, scenario, kind):
        if kind == nat.FG_SCENARIO_HD:
            self.h_shape[0] = np.asarray(scenario.ideal_shape, np.float64)
            self.h_ivel[0] = np.asarray(scenario.ideal_vel, np.float64)
        if self.L > 0:
            self.h_lm[0] = np.stack([np.asarray(l.state.p_pos, np.float64) for l in world.landmarks])
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self.h_lmv[0] = np.stack([np.zeros(2) if l.state.p_vel is None else np.asarray(l.state.p_vel, np.float64)
                                      for l in world.landmarks])

    def _scatter_obstacles(self, world, with_pos):
        """Obstacle state back to the host records: positions after a fused step, velocities as the reward
        hook's rule leaves them (formation_hd_obs_env.py:85-88)."""
        n = self._n_obst
        if not n:
            return
        lm = self.h_lm[0].astype(np.float64)
        lv = self.h_lmv[0].astype(np.float64)
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
        p, _ = self._params(world, kind, False, scenario)      # cached on every constant the kernels read
        key = (P.tobytes(), V.tobytes(), Cm.tobytes(), lmh.tobytes(),
               np.asarray(getattr(scenario, "ideal_shape", 0.0), np.float64).tobytes(),
               np.asarray(getattr(scenario, "ideal_vel", 0.0), np.float64).tobytes(),
               self._pkey)                                      # kind, sizes, collide, num_obs, obs_range, ...
        if key == self._cache_key:
            if kind == nat.FG_SCENARIO_HD_OBSTACLE:
                self._scatter_obstacles(world, with_pos=False)
            return self._cache_val
        b = self._buffers(kind, True, scenario)
        self.h_pos[0], self.h_vel[0], self.h_comm[0] = P, V, Cm
        self._stage_scenario(world, scenario, kind)
        fn = getattr(self.lib, "fg_obs_reward" + self.sfx)
        with torch.cuda.device(self.device):
            self._h2d()
            nat.check(fn(C.byref(p), C.byref(b), kind, 1, self.N, self.L, self._stream()), "fg_obs_reward")
            self._d2h("obs")
        self.launches += 1
        val = dict(obs=self._h_obs().astype(np.float64), indiv=self.h_indiv[0].astype(np.float64),
                   reward=float(self.h_reward[0, 0, 0]))
        if kind == nat.FG_SCENARIO_HD and self.L > 0:
            # observation's side effect (formation_hd_env.py:40-44): landmarks re-centred
            lm = self.h_lm[0].astype(np.float64)
            for k, l in enumerate(world.landmarks):
                l.state.p_pos = lm[k].copy()
            key = key[:3] + (lm.tobytes(),) + key[4:]
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self._scatter_obstacles(world, with_pos=False)       # the reward hook's velocity rule
        self._cache_key, self._cache_val = key, val
        return val

    # ------------------------------------------------------------------ fused env.step
    def step_fused(self, world, scenario, kind, acts, current_step, acts_c=None):
        """environment.py:113-142 in one launch.  ``current_step`` is the value BEFORE the step."""
        p, silent = self._params(world, kind, False, scenario)
        b = self._buffers(kind, True, scenario)
        self._stage_state(world)
        self._stage_actions(np.asarray(acts, np.float64), silent, acts_c)
        self._stage_scenario(world, scenario, kind)
        self.h_step[0] = int(current_step)
        fn = getattr(self.lib, "fg_step_fused" + self.sfx)
        with torch.cuda.device(self.device):
            self._h2d()
            nat.check(fn(C.byref(p), C.byref(b), kind, 1, self.N, self.L, 1, 0, 0,
                         int(world.seed) & (2 ** 64 - 1), int(world.world_step) & 0xFFFFFFFF, 0,
                         self._stream()), "fg_step_fused")
            self._d2h("obs")
        self.launches += 1
        self._scatter_state(world)
        if kind == nat.FG_SCENARIO_HD and self.L > 0:
            lm = self.h_lm[0].astype(np.float64)
            for k, l in enumerate(world.landmarks):
                l.state.p_pos = lm[k].copy()
        if kind == nat.FG_SCENARIO_HD_OBSTACLE:
            self._scatter_obstacles(world, with_pos=True)
        self._cache_key = None
        return dict(obs=self._h_obs().astype(np.float64), indiv=self.h_indiv[0].astype(np.float64),
                    reward=float(self.h_reward[0, 0, 0]), done=bool(self.h_done[0, 0]))
