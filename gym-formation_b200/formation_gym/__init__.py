"""formation_gym -- B200-native drop-in for jc-bao/gym-formation's MPE step path.

    import formation_gym
    env = formation_gym.make_env('formation_hd_env', benchmark=False, num_agents=9, episode_length=25)
    obs_n = env.reset()
    obs_n, reward_n, done_n, info_n = env.step(act_n)

    venv = formation_gym.make_batched_env('formation_hd_env', num_envs=131072, num_agents=9)
    obs = venv.reset(); obs, rew, done, info = venv.step(actions)      # [E,N,*] CUDA tensors

All arithmetic of the step path runs in hand-written sm_100a kernels behind the C ABI in
``include/formation_gym_b200.h``; there is no CPU fallback (importing works without a GPU, any
call that needs the device raises).
"""
import importlib

__all__ = ["make_env", "make_batched_env", "make_vec_env", "MultiAgentEnv", "BatchedFormationEnv", "CudaVecEnv",
           "SCENARIOS"]

SCENARIOS = ("basic_formation_env", "formation_hd_env")


def _load_scenario(scenario_name):
    if scenario_name.endswith(".py"):
        scenario_name = scenario_name[:-3]
    if scenario_name not in SCENARIOS:
        raise ValueError("scenario %r is not part of the accelerated step path (supported: %s)"
                         % (scenario_name, ", ".join(SCENARIOS)))
    return importlib.import_module("formation_gym.envs." + scenario_name).Scenario()


def make_env(scenario_name='basic_formation_env', benchmark=False, num_agents=3, episode_length=None):
    """Reference signature (formation_gym/__init__.py:6) plus the README's 4th argument
    ``episode_length`` (README.md:61).  Returns a ``MultiAgentEnv``."""
    from .environment import MultiAgentEnv
    scenario = _load_scenario(scenario_name)
    if scenario_name == "formation_hd_env" and episode_length is not None:
        world = scenario.make_world(num_agents, episode_length)
    else:
        world = scenario.make_world(num_agents)
        if episode_length is not None:
            world.world_length = episode_length
    if benchmark:
        return MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation,
                             scenario.benchmark_data, shared_viewer=True)
    return MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation,
                         shared_viewer=True)


def make_batched_env(scenario_name='formation_hd_env', num_envs=4096, num_agents=9,
                     episode_length=None, **kwargs):
    """Batched front end: ``num_envs`` envs as ``[envs, agents, dim]`` device tensors."""
    from .batched import BatchedFormationEnv
    return BatchedFormationEnv(scenario_name, num_envs, num_agents, episode_length, **kwargs)


def make_vec_env(scenario_name='formation_hd_env', num_envs=128, num_agents=9, episode_length=None, **kwargs):
    """VecEnv-compatible adapter (``SubprocVecEnv`` / ``DummyVecEnv`` interface of the reference's trainers,
    train/maddpg-v2/utils/env_wrappers.py:40-128) over the batched CUDA env."""
    from .vec_env import CudaVecEnv
    return CudaVecEnv(scenario_name, num_envs, num_agents, episode_length, **kwargs)


def __getattr__(name):
    if name == "CudaVecEnv":
        from .vec_env import CudaVecEnv
        return CudaVecEnv
    if name == "MultiAgentEnv":
        from .environment import MultiAgentEnv
        return MultiAgentEnv
    if name == "BatchedFormationEnv":
        from .batched import BatchedFormationEnv
        return BatchedFormationEnv
    raise AttributeError(name)
