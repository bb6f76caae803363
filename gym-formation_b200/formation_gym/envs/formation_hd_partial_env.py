"""formation_hd_partial_env scenario: partial observation -- every agent sees the landmarks (absolute
positions) and only the next ``num_obs`` agents (cyclic); reward = -Hausdorff(agents - mean,
landmarks - mean) - #collisions with threshold s1 + s2 (reference:
formation_gym/envs/formation_hd_partial_env.py:15-125).

Host hooks build / initialise the world like the reference (same ``np.random`` draw order: agents, then
landmarks); ``observation`` / ``reward`` come from the sm_100a kernel (``fg_obs_reward``, scenario id
FG_SCENARIO_HD_PARTIAL) and ``MultiAgentEnv.step`` uses the fused step kernel."""
import numpy as np

from .. import _native as nat
from ..core import World, Agent, Landmark
from ..scenario import BaseScenario


class Scenario(BaseScenario):
    native_kind = nat.FG_SCENARIO_HD_PARTIAL

    def make_world(self, num_agents=5, num_landmarks=5, num_obs=3, world_length=25):
        self.num_obs = num_obs
        self.num_agents = num_agents
        return self._build(num_agents, num_landmarks, world_length)

    def _build(self, num_agents, num_landmarks, world_length):
        world = World()
        world.world_length = world_length
        world.dim_c = 2
        world.collaborative = True
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.04
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.02
        self.reset_world(world)
        return world

    def reset_world(self, world):
        """Initial conditions (host hook; draw order of formation_hd_partial_env.py:89-99)."""
        for agent in world.agents:
            agent.color = np.array([0.35, 0.35, 0.85])
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.color = np.array([0.25, 0.25, 0.25])
            landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def _eval(self, world):
        return world.backend().scenario_eval(world, self, self.native_kind)

    @staticmethod
    def _index(agent, world):
        for i, a in enumerate(world.agents):
            if a is agent:
                return i
        raise ValueError("agent does not belong to this world")

    def observation(self, agent, world):
        """[p_vel, landmark positions (2L), other_pos of the next num_obs agents, comm (2(N-1))]."""
        return self._eval(world)["obs"][self._index(agent, world)].copy()

    def reward(self, agent, world):
        return float(self._eval(world)["indiv"][self._index(agent, world)])

    def is_collision(self, agent1, agent2):
        d = agent1.state.p_pos - agent2.state.p_pos
        return float(np.sqrt(np.sum(np.square(d)))) < (agent1.size + agent2.size)

    def benchmark_data(self, agent, world):
        from .._bench_info import benchmark_info
        return benchmark_info(self, agent, world, half_threshold=False)

    def set_bound(self, world):
        """formation_hd_partial_env.py:127-129 (unused by the reference's step path)."""
        for agent in world.agents:
            agent.state.p_pos = np.clip(agent.state.p_pos, [-2, -2], [2, 2])
