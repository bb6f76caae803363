"""``MultiAgentEnv`` -- the single-env facade with the reference's API
(formation_gym/environment.py:11-236 of jc-bao/gym-formation), backed by the sm_100a kernels.

``step(action_n) -> (obs_n, reward_n, done_n, info_n)`` keeps the reference's types: lists of
float64 ndarrays, ``[[R]] * N`` (one aliased inner list, environment.py:138), python bools and
``{'individual_reward': r_i}`` dicts.  With the stock scenario hooks the whole step --
``_set_action``, ``World.step``, every agent's observation and reward, done, the shared-reward
sum -- is ONE fused kernel launch (``fg_step_fused_f64``, E = 1).  With user-supplied callbacks
the physics still runs on the GPU (``World.step``) and the callbacks are called like the
reference does.

Deliberate differences: the caller's action arrays are not scaled in
place (Q4), the reward is evaluated once per agent instead of twice (Q5), ``make_env`` accepts
``episode_length`` (Q1) and rendering (pyglet) is not part of this package.
"""
import numpy as np

from . import spaces


class MultiAgentEnv(object):
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, world, reset_callback=None, reward_callback=None,
                 observation_callback=None, info_callback=None,
                 done_callback=None, post_step_callback=None,
                 shared_viewer=True, discrete_action=False):
        self.world = world
        self.world_length = self.world.world_length
        self.current_step = 0
        self.agents = self.world.policy_agents
        self.num_agents = len(world.policy_agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.post_step_callback = post_step_callback
        self.num_envs = 1
        self.discrete_action_space = discrete_action
        self.discrete_action_input = False
        self.force_discrete_action = getattr(world, 'discrete_action', False)
        self.shared_reward = getattr(world, 'collaborative', False)
        self.time = 0

        self.action_space = []
        self.observation_space = []
        share_obs_dim = 0
        for agent in self.agents:
            total = []
            if self.discrete_action_space:
                u_space = spaces.Discrete(world.dim_p * 2 + 1)
            else:
                u_space = spaces.Box(low=-agent.u_range, high=+agent.u_range,
                                     shape=(world.dim_p,), dtype=np.float32)
            if agent.movable:
                total.append(u_space)
            if self.discrete_action_space:
                c_space = spaces.Discrete(world.dim_c)
            else:
                c_space = spaces.Box(low=0.0, high=1.0, shape=(world.dim_c,), dtype=np.float32)
            if not agent.silent:
                total.append(c_space)
            self.action_space.append(spaces.Tuple(total) if len(total) > 1 else total[0])
            obs_dim = self._obs_dim(agent)
            share_obs_dim += obs_dim
            self.observation_space.append(spaces.Box(low=-np.inf, high=+np.inf, shape=(obs_dim,),
                                                     dtype=np.float32))
            agent.action.c = np.zeros(self.world.dim_c)
        self.share_observation_space = [
            spaces.Box(low=-np.inf, high=+np.inf, shape=(share_obs_dim,), dtype=np.float32)
            for _ in range(self.num_agents)]
        self.shared_viewer = shared_viewer
        self.viewers = [None] if shared_viewer else [None] * self.num_agents

    # ------------------------------------------------------------------ helpers
    def _obs_dim(self, agent):
        """Observation length.  The reference calls the observation hook once per agent just to
        measure it (environment.py:89); the stock scenarios' lengths are known in closed form, so
        constructing an env needs no device call."""
        sc = self._native_scenario()
        if sc is not None:
            from .batched import obs_dim, SCENARIOS
            name = [k for k, v in SCENARIOS.items() if v == sc.native_kind][0]
            return obs_dim(name, len(self.world.agents), len(self.world.landmarks),
                           int(getattr(sc, "num_obs", 3) or 0))
        return len(self.observation_callback(agent, self.world))

    def seed(self, seed=None):
        """Seeds the host RNG of the scenario hooks like the reference (environment.py:106-110)
        and the Philox key of the device-side motor/comm noise."""
        s = 1 if seed is None else seed
        np.random.seed(s)
        self.world.seed = int(s)

    def _native_scenario(self):
        """The stock Scenario behind the callbacks, or None when the hooks were replaced."""
        obs_cb, rew_cb = self.observation_callback, self.reward_callback
        sc = getattr(obs_cb, '__self__', None)
        if sc is None or getattr(rew_cb, '__self__', None) is not sc:
            return None
        kind = getattr(sc, 'native_kind', None)
        if kind is None:
            return None
        cls = type(sc)
        from .envs import (formation_hd_env, basic_formation_env, formation_hd_partial_env,
                           formation_hd_partial_range_env, formation_hd_obs_env)
        # most-derived stock class first: formation_hd_partial_range_env.Scenario subclasses the partial one
        # (and overrides its hooks), so a plain isinstance() walk would match the wrong stock class
        stocks = (formation_hd_partial_range_env.Scenario, formation_hd_partial_env.Scenario,
                  formation_hd_obs_env.Scenario, formation_hd_env.Scenario, basic_formation_env.Scenario)
        stock = next((s for s in stocks if cls is s), None) or next((s for s in stocks if isinstance(sc, s)), None)
        if stock is None or stock.native_kind != kind:
            return None
        # a user subclass that overrides a hook gets the callback path (its Python code must run)
        if cls.observation is not stock.observation or cls.reward is not stock.reward:
            return None
        return sc

    def _decode_u(self, action, agent):
        """The physical part of ``_set_action`` BEFORE the sensitivity scaling
        (environment.py:192-217): continuous, one-hot discrete or integer actions -> u."""
        if self.discrete_action_input:
            u = np.zeros(self.world.dim_p)
            if action == 1: u[0] = -1.0
            if action == 2: u[0] = +1.0
            if action == 3: u[1] = -1.0
            if action == 4: u[1] = +1.0
            return u
        a = np.array(action, dtype=np.float64)          # a copy: the caller's array is untouched
        if self.discrete_action_space:
            return np.array([a[1] - a[2], a[3] - a[4]])
        if self.force_discrete_action:
            p = int(np.argmax(a[0:self.world.dim_p]))
            a[:] = 0.0
            a[p] = 1.0
        return a[0:self.world.dim_p]

    def _split_action(self, action, agent):
        """Split one agent's action into (physical, comm) parts like ``action = [action]`` +
        ``action[1:]`` bookkeeping in the reference (environment.py:191,223,233)."""
        if agent.movable and not agent.silent:
            return action[0], action[1]
        if agent.movable:
            return action, None
        return None, action

    def _set_action(self, action, agent, action_space=None, time=None):
        agent.action.u = np.zeros(self.world.dim_p)
        agent.action.c = np.zeros(self.world.dim_c)
        a_u, a_c = self._split_action(action, agent)
        if agent.movable:
            sensitivity = 5.0 if agent.accel is None else agent.accel
            agent.action.u = self._decode_u(a_u, agent) * sensitivity
        if not agent.silent:
            if self.discrete_action_input:
                agent.action.c = np.zeros(self.world.dim_c)
                agent.action.c[a_c] = 1.0
            else:
                agent.action.c = np.array(a_c, dtype=np.float64)

    def _get_info(self, agent):
        return {} if self.info_callback is None else self.info_callback(agent, self.world)

    def _get_obs(self, agent):
        if self.observation_callback is None:
            return np.zeros(0)
        return self.observation_callback(agent, self.world)

    def _get_done(self, agent):
        if self.done_callback is None:
            return self.current_step >= self.world_length
        return self.done_callback(agent, self.world)

    def _get_reward(self, agent):
        return 0.0 if self.reward_callback is None else self.reward_callback(agent, self.world)

    # ------------------------------------------------------------------ gym API
    def step(self, action_n):
        self.agents = self.world.policy_agents
        sc = self._native_scenario()
        # (World.cache_dists keeps World.step() a separate call: the cache is refreshed between the physics and the hooks)
        fused = (sc is not None and self.done_callback is None and not self.world.scripted_agents
                 and all(a.movable for a in self.agents) and not getattr(self.world, "cache_dists", False))
        self.current_step += 1
        obs_n, reward_n, done_n, info_n = [], [], [], []
        if fused:
            silent = all(a.silent for a in self.agents)
            n, dim_p = len(self.agents), self.world.dim_p
            acts_c = []
            plain = silent and not (self.discrete_action_input or self.discrete_action_space
                                    or self.force_discrete_action)
            U = None
            if plain:
                try:                                        # the common case in one conversion: N arrays of dim_p floats
                    U = np.array(action_n, dtype=np.float64)
                    if U.shape != (n, dim_p):
                        U = U[:, :dim_p] if (U.ndim == 2 and U.shape[0] == n and U.shape[1] >= dim_p) else None
                except (ValueError, TypeError):
                    U = None
            if U is None:
                U = np.empty((n, dim_p))
                for i, agent in enumerate(self.agents):
                    a_u, a_c = self._split_action(action_n[i], agent)
                    U[i] = self._decode_u(a_u, agent)
                    if not silent:
                        acts_c.append(np.array(a_c, dtype=np.float64))
            # host records as _set_action leaves them (environment.py:188-236): callbacks and render code that read
            # agent.action see the same values on the fused and on the callback path
            zc = np.zeros(self.world.dim_c)
            for i, agent in enumerate(self.agents):
                agent.action.u = U[i] * (5.0 if agent.accel is None else agent.accel)
                agent.action.c = zc.copy() if silent else acts_c[i].copy()
            self.world.world_step += 1
            out = self.world.backend().step_fused(
                self.world, sc, sc.native_kind, U, self.current_step - 1,
                None if silent else np.stack(acts_c))
            obs_n = list(out["obs"])                        # rows of a fresh array (nothing aliases the backend)
            indiv = out["indiv"].tolist()
            done = bool(out["done"])
            for i, agent in enumerate(self.agents):
                reward_n.append([indiv[i]])
                done_n.append(done)
                info = {'individual_reward': indiv[i]}
                if self.info_callback is not None:
                    env_info = self._get_info(agent)
                    if 'fail' in env_info.keys():
                        info['fail'] = env_info['fail']
                info_n.append(info)
            reward = out["reward"]
        else:
            for i, agent in enumerate(self.agents):
                self._set_action(action_n[i], agent, self.action_space[i])
            self.world.step()
            for i, agent in enumerate(self.agents):
                obs_n.append(self._get_obs(agent))
                r = self._get_reward(agent)
                reward_n.append([r])
                done_n.append(self._get_done(agent))
                info = {'individual_reward': r}
                env_info = self._get_info(agent)
                if 'fail' in env_info.keys():
                    info['fail'] = env_info['fail']
                info_n.append(info)
            reward = np.sum(reward_n)
        if self.shared_reward:
            reward_n = [[reward]] * self.num_agents
        if self.post_step_callback is not None:
            self.post_step_callback(self.world)
        return obs_n, reward_n, done_n, info_n

    def reset(self):
        self.current_step = 0
        self.reset_callback(self.world)
        self.agents = self.world.policy_agents
        return [self._get_obs(agent) for agent in self.agents]

    def render(self, mode='human', close=False):
        """Reference signature (environment.py:243-393).  Host-side visualisation of the facade's records through
        ``formation_gym.render_bridge`` (numpy rasteriser; the reference's pyglet viewer is not needed):
        ``mode='rgb_array'`` returns one 700 x 700 x 3 uint8 frame per viewer, ``mode='human'`` shows it in a pyglet
        window when pyglet is importable."""
        from . import render_bridge as rb
        if close:
            for i, viewer in enumerate(self.viewers):
                if viewer is not None and hasattr(viewer, "close"):
                    viewer.close()
                self.viewers[i] = None
            return []
        results = []
        for i in range(len(self.viewers)):
            center = None if self.shared_viewer else self.agents[i].state.p_pos        # :363-367
            frame = rb.world_frame(self.world, self.agents, center)
            if mode == 'rgb_array':
                results.append(frame)
            else:
                self.viewers[i] = rb.show(frame, self.viewers[i])
                results.append(True)
        return results
