"""formation_hd_env scenario: N agents, N landmarks that define the ideal formation; reward =
-Hausdorff(shape) - |ideal_vel - mean_vel| - #collisions (reference:
formation_gym/envs/formation_hd_env.py:13-121).

``make_world`` / ``reset_world`` build and initialise host-side records exactly like the
reference (same RNG draw order on ``np.random``: agents, landmarks, ideal_vel -- so a seeded reset
reproduces the reference's initial state bit for bit).  ``observation`` and ``reward`` are
evaluated for all agents at once by the sm_100a kernel (``fg_obs_reward``) and served from that
result; ``MultiAgentEnv.step`` bypasses the per-agent hooks entirely and uses the fused step kernel.
"""
import numpy as np

from .. import _native as nat
from ..core import World, Agent, Landmark
from ..scenario import BaseScenario


class Scenario(BaseScenario):
    native_kind = nat.FG_SCENARIO_HD

    def make_world(self, num_agents=3, episode_length=100):
        if num_agents < 3:
            raise ValueError("formation_hd_env needs num_agents >= 3")
        world = World()
        world.world_length = episode_length
        world.dim_c = 2
        world.collaborative = True
        self.num_agents = num_agents
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.03
        world.landmarks = [Landmark() for _ in range(num_agents)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmarks %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.01
        self.reset_world(world)
        return world

    def reset_world(self, world):
        """Initial conditions (host hook; draw order of formation_hd_env.py:77-95)."""
        for agent in world.agents:
            agent.color = np.array([0.35, 0.35, 0.85])
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        raw = []
        for landmark in world.landmarks:
            landmark.color = np.array([0.25, 0.25, 0.25])
            raw.append(np.random.uniform(-1, +1, world.dim_p))
            landmark.state.p_pos = raw[-1]
            landmark.state.p_vel = np.zeros(world.dim_p)
        self.ideal_shape = raw - np.mean(raw, 0)
        self.ideal_vel = np.random.uniform(-1, +1, world.dim_p)

    def _eval(self, world):
        return world.backend().scenario_eval(world, self, self.native_kind)

    @staticmethod
    def _index(agent, world):
        for i, a in enumerate(world.agents):
            if a is agent:
                return i
        raise ValueError("agent does not belong to this world")

    def observation(self, agent, world):
        """[p_vel, other_pos (2(N-1)), comm (2(N-1)), ideal_shape (2N), ideal_vel (2)] = 6N."""
        return self._eval(world)["obs"][self._index(agent, world)].copy()

    def reward(self, agent, world):
        return float(self._eval(world)["indiv"][self._index(agent, world)])

    def is_collision(self, agent1, agent2):
        d = agent1.state.p_pos - agent2.state.p_pos
        return float(np.sqrt(np.sum(np.square(d)))) < (agent1.size + agent2.size) / 2

    def benchmark_data(self, agent, world):
        from .._bench_info import benchmark_info
        return benchmark_info(self, agent, world, half_threshold=True)

    def generate_shape(self, layer, layer_shapes=None):
        """Hierarchical target shapes (formation_hd_env.py:123-139): host-side configuration data."""
        if layer_shapes is None:
            layer_shapes = np.array([
                [[0, -1], [0.5, 0], [0, 1]],
                [[0, 1.6], [-1, 0], [1, 0]],
                [[1.5, 0], [0, 0], [-1.5, 0]],
                [[0, 0.6], [1, 0], [-1, 0]],
            ])
        assert layer < layer_shapes.shape[0], 'Layer shape is not enough!'
        if layer == 0:
            return layer_shapes[0]
        inner = self.generate_shape(layer - 1)
        return np.array([layer_shapes[layer][i] + inner * 0.45 for i in range(layer_shapes.shape[1])])
