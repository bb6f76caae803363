"""formation_hd_partial_range_env scenario: like formation_hd_partial_env, but every agent sees ALL others
with the relative position clipped to [-obs_range, obs_range] per axis (reference:
formation_gym/envs/formation_hd_partial_range_env.py:15-117).  Scenario id FG_SCENARIO_HD_PARTIAL_RANGE."""
from .. import _native as nat
from . import formation_hd_partial_env as _partial


class Scenario(_partial.Scenario):
    native_kind = nat.FG_SCENARIO_HD_PARTIAL_RANGE
    num_obs = 0

    def make_world(self, num_agents=4, num_landmarks=4, obs_range=0.7, world_length=25):
        self.obs_range = obs_range
        self.num_agents = num_agents
        return self._build(num_agents, num_landmarks, world_length)

    def observation(self, agent, world):
        """[p_vel, landmark positions (2L), clip(other_pos, +-obs_range) (2(N-1)), comm (2(N-1))]."""
        return self._eval(world)["obs"][self._index(agent, world)].copy()
