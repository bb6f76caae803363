"""formation_hd_obs_env scenario: formation control among falling obstacles -- ``num_landmarks`` goal landmarks
(immovable, non-colliding) and ``num_obstacles`` obstacle landmarks that are MOVABLE and COLLIDE with the agents and
with each other; the reward hook resets every obstacle's velocity to (0, -1) while it is above y = -2.2
(reference: formation_gym/envs/formation_hd_obs_env.py:14-149).

Host hooks build / initialise the world like the reference (same ``np.random`` draw order: agents, goal landmarks,
obstacles); ``observation`` / ``reward`` come from the sm_100a kernel (``fg_obs_reward``, scenario id
FG_SCENARIO_HD_OBSTACLE) and ``MultiAgentEnv.step`` uses the fused step kernel (``fg_step_fused``), which also
integrates the obstacles and applies the velocity rule."""
import numpy as np

from .. import _native as nat
from ..core import World, Agent, Landmark
from ..scenario import BaseScenario


class Scenario(BaseScenario):
    native_kind = nat.FG_SCENARIO_HD_OBSTACLE
    num_obs = 0

    def make_world(self, num_agents=4, num_landmarks=4, num_obstacles=3, world_length=50):
        self.num_agents = num_agents
        self.num_landmarks = num_landmarks
        self.num_obstacles = num_obstacles
        world = World()
        world.world_length = world_length
        world.dim_c = 2
        world.collaborative = True
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.1
        world.landmarks = [Landmark() for _ in range(num_landmarks + num_obstacles)]
        for i, landmark in enumerate(world.landmarks):
            if i < num_landmarks:
                landmark.name = 'landmarks %d' % i
                landmark.collide = False
                landmark.movable = False
                landmark.size = 0.02
            else:
                landmark.name = 'obstacles %d' % (i - num_landmarks)
                landmark.collide = True
                landmark.movable = True
                landmark.size = 0.15
        self.reset_world(world)
        return world

    def reset_world(self, world):
        """Initial conditions (host hook; draw order of formation_hd_obs_env.py:101-120)."""
        for agent in world.agents:
            agent.color = np.array([0.65, 0.65, 0.85])
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        step = np.linspace(-1.8, 1.8, self.num_obstacles + 1)
        for i, landmark in enumerate(world.landmarks):
            if i < self.num_landmarks:
                landmark.color = np.array([0, 0.6, 0])
                landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
                landmark.state.p_vel = np.zeros(world.dim_p)
            else:
                k = i - self.num_landmarks
                landmark.color = np.array([0.25, 0.25, 0.25])
                landmark.state.p_pos = np.random.uniform([step[k], 2.0], [step[k + 1], 2.5])
                landmark.state.p_vel = np.array([0.0, -1.0])

    def _eval(self, world):
        return world.backend().scenario_eval(world, self, self.native_kind)

    @staticmethod
    def _index(agent, world):
        for i, a in enumerate(world.agents):
            if a is agent:
                return i
        raise ValueError("agent does not belong to this world")

    def observation(self, agent, world):
        """[p_vel, goal landmark positions, obstacle positions - p_i, other_pos (2(N-1)), comm (2(N-1))]."""
        return self._eval(world)["obs"][self._index(agent, world)].copy()

    def reward(self, agent, world):
        """-Hausdorff(agents - mean, goals - mean) - 2 per collision with another agent or an obstacle; like
        the reference's hook it also rewrites the obstacles' velocities (done by the kernel)."""
        return float(self._eval(world)["indiv"][self._index(agent, world)])

    def is_collision(self, agent1, agent2):
        d = agent1.state.p_pos - agent2.state.p_pos
        return float(np.sqrt(np.sum(np.square(d)))) < (agent1.size + agent2.size)

    def benchmark_data(self, agent, world):
        from .._bench_info import benchmark_info
        return benchmark_info(self, agent, world, half_threshold=False)

    def set_bound(self, world):
        """formation_hd_obs_env.py:151-155 (unused by the reference's step path)."""
        for agent in world.agents:
            agent.state.p_pos = np.clip(agent.state.p_pos, [-2.5, -20], [2.5, 20])
        for landmark in world.landmarks[self.num_landmarks:]:
            landmark.state.p_pos = np.clip(landmark.state.p_pos, [-2.5, -20], [2.5, 20])
