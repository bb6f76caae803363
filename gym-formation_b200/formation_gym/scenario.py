"""Scenario plugin base class (same hooks as formation_gym/scenario.py:4-12 of the reference)."""


class BaseScenario(object):
    """A scenario builds a ``World`` and supplies the reset / observation / reward hooks that
    ``MultiAgentEnv`` calls (formation_gym/environment.py:16-19)."""

    def make_world(self):
        raise NotImplementedError()

    def reset_world(self, world):
        raise NotImplementedError()

    def info(self, agent, world):
        return {}
