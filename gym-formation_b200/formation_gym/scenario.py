"""Scenario plugin interface (the hook names of formation_gym/scenario.py:4-12 of the reference).

A scenario builds a ``core.World`` and supplies the reset / observation / reward hooks that ``MultiAgentEnv`` calls
(formation_gym/environment.py:16-19).  Stock scenarios additionally carry ``native_kind``, the FG_SCENARIO_* id under
which the CUDA library evaluates their ``observation`` / ``reward`` hooks and the fused step."""


class BaseScenario(object):
    native_kind = None          # set by the stock scenarios in formation_gym.envs

    def _missing(self, hook, what):
        return NotImplementedError("%s.%s is not implemented: %s" % (type(self).__name__, hook, what))

    def make_world(self, *args, **kwargs):
        raise self._missing("make_world", "build the agents / landmarks and return a formation_gym.core.World")

    def reset_world(self, world):
        raise self._missing("reset_world", "set the initial conditions of every entity of `world`")

    def info(self, agent, world):
        """Optional per-agent diagnostics; nothing by default."""
        return dict()
