"""World / entity data model with the reference's attribute names (formation_gym/core.py:4-139)
and a ``World.step()`` that runs on the GPU.

The classes are plain host-side records exactly like the reference's (user code reads and writes
``agent.state.p_pos`` etc.); ``World.step()`` gathers them into ``[1,N,2]`` device tensors, calls
the fused sm_100a physics kernel through the C ABI (``fg_world_step``: apply_action_force,
apply_environment_force / get_entity_collision_force, wall forces, integrate_state,
update_agent_state -- formation_gym/core.py:206-362) and scatters the result back.  There is no
Python/numpy implementation of the physics in this package.
"""
import numpy as np

# Bumped whenever a constant the kernels read (sizes, masses, flags, walls, world constants ...) is assigned on a
# World / Entity / Wall record, so that the device backend rebuilds its parameter block only then instead of
# comparing every attribute of every agent on every step.  Per-step records (state, action, counters) do not count.
_config_version = [0]
_PER_STEP_ATTRS = frozenset(("state", "action", "world_step", "cached_dist_vect", "cached_dist_mag", "_backend",
                             "min_dists", "cached_collisions",
                             "seed", "color", "name", "i", "channel", "goal"))


def config_version():
    return _config_version[0]


class _Tracked(object):
    def __setattr__(self, name, value):
        if name not in _PER_STEP_ATTRS:
            _config_version[0] += 1
        object.__setattr__(self, name, value)


class EntityState(object):
    def __init__(self):
        self.p_pos = None      # physical position
        self.p_vel = None      # physical velocity


class AgentState(EntityState):
    def __init__(self):
        super(AgentState, self).__init__()
        self.c = None          # communication utterance


class Action(object):
    def __init__(self):
        self.u = None          # physical action
        self.c = None          # communication action


class Wall(_Tracked):
    def __init__(self, orient='H', axis_pos=0.0, endpoints=(-1, 1), width=0.1, hard=True):
        self.orient = orient                   # 'H'orizontal or 'V'ertical
        self.axis_pos = axis_pos               # y for H, x for V
        self.endpoints = np.array(endpoints)   # extent along the wall
        self.width = width
        self.hard = hard                       # impassable to all agents
        self.color = np.array([0.0, 0.0, 0.0])


class Entity(_Tracked):
    def __init__(self):
        self.i = 0
        self.name = ''
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.ghost = False
        self.density = 25.0
        self.color = None
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0
        self.channel = None

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    pass


class Agent(Entity):
    def __init__(self):
        super(Agent, self).__init__()
        self.adversary = False
        self.dummy = False
        self.movable = True
        self.silent = False
        self.blind = False
        self.u_noise = None
        self.c_noise = None
        self.u_range = 1.0
        self.state = AgentState()
        self.action = Action()
        self.action_callback = None
        self.goal = None


class World(_Tracked):
    """Multi-agent world.  Same public attributes as the reference (core.py:112-139)."""

    def __init__(self, world_length=50):
        self.agents = []
        self.landmarks = []
        self.walls = []
        self.dim_c = 0
        self.dim_p = 2
        self.dim_color = 3
        self.dt = 0.1
        self.damping = 0.25
        self.contact_force = 1e+2
        self.contact_margin = 1e-3
        self.cache_dists = False
        self.cached_dist_vect = None
        self.cached_dist_mag = None
        self.min_dists = None
        self.cached_collisions = None
        self.world_length = world_length
        self.world_step = 0
        self.num_agents = 0
        self.num_landmarks = 0
        # B200 backend (created lazily on the first device call)
        self.dtype = np.float64          # facade precision: fp64 kernels (the reference is fp64)
        self.seed = 1                    # Philox key for u_noise / c_noise (env.seed() sets it)
        self._backend = None

    @property
    def entities(self):
        return self.agents + self.landmarks

    @property
    def policy_agents(self):
        return [agent for agent in self.agents if agent.action_callback is None]

    @property
    def scripted_agents(self):
        return [agent for agent in self.agents if agent.action_callback is not None]

    def assign_agent_colors(self):
        n_dummies = len([a for a in self.agents if getattr(a, 'dummy', False)])
        n_adv = len([a for a in self.agents if getattr(a, 'adversary', False)])
        n_good = len(self.agents) - n_adv - n_dummies
        colors = [(0.25, 0.75, 0.25)] * n_dummies + [(0.75, 0.25, 0.25)] * n_adv + \
                 [(0.25, 0.25, 0.75)] * n_good
        for color, agent in zip(colors, self.agents):
            agent.color = color

    def assign_landmark_colors(self):
        for landmark in self.landmarks:
            landmark.color = np.array([0.25, 0.25, 0.25])

    def backend(self):
        from ._facade import FacadeBackend
        if self._backend is None or not self._backend.matches(self):
            self._backend = FacadeBackend(self)
        return self._backend

    def step(self):
        """Advance the world by one step on the GPU (reference: core.py:206-225).

        ``agent.action.u`` is taken as already scaled by ``MultiAgentEnv._set_action``
        (environment.py:216-221), exactly as the reference's ``World.step`` expects."""
        self.world_step += 1
        for agent in self.scripted_agents:
            agent.action = agent.action_callback(agent, self)
        self.backend().world_step(self)
        if self.cache_dists:                       # core.py:224-225
            self.calculate_distances()

    def calculate_distances(self):
        """Reference signature (core.py:156-180): the distance cache behind ``cache_dists`` -- ``cached_dist_vect``,
        ``cached_dist_mag``, ``min_dists``, ``cached_collisions`` over ``self.entities`` -- computed on the GPU."""
        self.backend().calculate_distances(self)
