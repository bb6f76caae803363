"""basic_formation_env scenario (simple_spread style): reward = -sum_l min_a |p_a - l| minus
collisions, self included (reference: formation_gym/envs/basic_formation_env.py:7-91).

Host hooks build / initialise the world like the reference; ``observation`` / ``reward`` come from
the sm_100a kernel (``fg_obs_reward``), ``MultiAgentEnv.step`` uses the fused step kernel."""
import numpy as np

from .. import _native as nat
from ..core import World, Agent, Landmark
from ..scenario import BaseScenario


class Scenario(BaseScenario):
    native_kind = nat.FG_SCENARIO_BASIC

    def make_world(self, num_agents=3, num_landmarks=3):
        world = World()
        world.dim_c = 2
        world.collaborative = True
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.1
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmarks %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        """Initial conditions (host hook; draw order of basic_formation_env.py:54-65)."""
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
        """[p_vel, p_pos, landmark - p (2L), other_pos (2(N-1)), comm (2(N-1))]."""
        return self._eval(world)["obs"][self._index(agent, world)].copy()

    def reward(self, agent, world):
        return float(self._eval(world)["indiv"][self._index(agent, world)])

    def is_collision(self, agent1, agent2):
        d = agent1.state.p_pos - agent2.state.p_pos
        return float(np.sqrt(np.sum(np.square(d)))) < (agent1.size + agent2.size)

    def benchmark_data(self, agent, world):
        from .._bench_info import benchmark_info
        return benchmark_info(self, agent, world, half_threshold=False)
