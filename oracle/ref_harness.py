"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (jc-bao/gym-formation).

Nothing in the product path (``gym-formation_b200/``) may import this file.  It exists so that
``tests/golden/make_golden.py`` and the pinning tests can run the reference's own Python files
from ``/root/reference`` in THIS container (the GPU box has no ``/root/reference``).

The reference does not import as shipped on Python 3.12:
  * ``formation_gym/__init__.py:1``   imports ``imp`` (removed in 3.12)
  * ``formation_gym/environment.py:1-3`` imports ``gym`` (not installed)
  * ``formation_gym/envs/basic_formation_env.py:3-4`` imports the un-vendored ``multiagent``
So this module installs ~40 lines of in-process stub modules (``imp.load_source``, the handful
of ``gym`` names used for shapes/dtypes only, and ``multiagent`` aliased to the reference's own
``formation_gym.core`` / ``formation_gym.scenario`` -- see SURVEY.md 8c for why that alias is
the only faithful one) and then imports the reference package *unchanged* from disk.

Because the reference's import name (``formation_gym``) is the same as the drop-in package this
repo ships, this harness must only ever be used in a process that has NOT imported the repo's
own ``formation_gym`` (the tests run it through ``python oracle/ref_harness.py ...`` /
``make_golden.py`` subprocesses).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")           # written by oracle/make_ref.py (git-ignored; travels to the GPU box)


def _pick_root():
    """FG_REFERENCE_ROOT, else the read-only checkout (/root/reference, this container only), else the staged
    byte-for-byte copy under oracle/_ref (the GPU box), verified against its sha256 manifest."""
    env = os.environ.get("FG_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/formation_gym/core.py"):
        return "/root/reference"
    try:
        sys.path.insert(0, _HERE)
        import make_ref
        if make_ref.staged_ok(STAGED_ROOT):
            return STAGED_ROOT
    except Exception:
        pass
    finally:
        if sys.path and sys.path[0] == _HERE:
            sys.path.pop(0)
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "formation_gym", "core.py"))


def _install_stubs():
    # --- imp.load_source (formation_gym/__init__.py:9) -------------------------------------
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")

        def load_source(name, pathname):
            loader = importlib.machinery.SourceFileLoader(name or "_fg_scenario", pathname)
            spec = importlib.util.spec_from_loader(loader.name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            return mod

        imp.load_source = load_source
        sys.modules["imp"] = imp

    # --- gym (environment.py:1-3,65-96; multi_discrete.py:6,9): shapes/dtypes only ----------
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")

        class Env(object):
            pass

        class Space(object):
            pass

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

            def sample(self):
                return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)

        class Discrete(Space):
            def __init__(self, n):
                self.n = n
                self.shape = ()

        class Tuple(Space):
            def __init__(self, spaces_):
                self.spaces = tuple(spaces_)

        class EnvSpec(object):
            def __init__(self, *a, **k):
                pass

        gym.Env, gym.Space, gym.spaces, gym.envs = Env, Space, spaces, envs
        spaces.Box, spaces.Discrete, spaces.Tuple, spaces.Space = Box, Discrete, Tuple, Space
        envs.registration = registration
        registration.EnvSpec = EnvSpec
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                            "gym.envs.registration": registration})


def load_reference():
    """Import the unmodified reference package and return the module ``formation_gym``."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    mine = sys.modules.get("formation_gym")
    if mine is not None and not getattr(mine, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("the repo's own formation_gym is already imported in this process; "
                           "run the reference harness in a separate process")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # multiagent.{core,scenario} -> the reference's own core/scenario
    import formation_gym.core as ref_core          # noqa: E402  (reference file, unmodified)
    import formation_gym.scenario as ref_scenario  # noqa: E402
    ma = types.ModuleType("multiagent")
    ma.core, ma.scenario = ref_core, ref_scenario
    sys.modules.setdefault("multiagent", ma)
    sys.modules.setdefault("multiagent.core", ref_core)
    sys.modules.setdefault("multiagent.scenario", ref_scenario)
    import formation_gym                           # noqa: E402
    return formation_gym


def make_reference_env(scenario, num_agents, episode_length=None):
    """``formation_gym.make_env(scenario, False, num_agents)`` on the unmodified reference.

    The reference's ``make_env`` has no ``episode_length`` argument
    (formation_gym/__init__.py:6 vs README.md:61), so it is patched onto the built env the way
    a user of the reference would have to: ``world.world_length`` and ``env.world_length``
    (environment.py:22,174)."""
    fg = load_reference()
    env = fg.make_env(scenario, False, num_agents)
    if episode_length is not None:
        env.world.world_length = episode_length
        env.world_length = episode_length
    return env


def scenario_of(env):
    """The Scenario instance behind an env (holds ideal_shape/ideal_vel, formation_hd_env.py:86-95)."""
    return env.reset_callback.__self__


def inject_state(env, pos, vel, ideal_shape=None, ideal_vel=None, landmarks=None, step=0):
    """Overwrite the reference env's state in place (float64 copies)."""
    world = env.world
    for i, a in enumerate(world.agents):
        a.state.p_pos = np.array(pos[i], dtype=np.float64)
        a.state.p_vel = np.array(vel[i], dtype=np.float64)
        a.state.c = np.zeros(world.dim_c)
    sc = scenario_of(env)
    if ideal_shape is not None:
        sc.ideal_shape = np.array(ideal_shape, dtype=np.float64)
    if ideal_vel is not None:
        sc.ideal_vel = np.array(ideal_vel, dtype=np.float64)
    if landmarks is not None:
        for k, l in enumerate(world.landmarks):
            l.state.p_pos = np.array(landmarks[k], dtype=np.float64)
            l.state.p_vel = np.zeros(world.dim_p)
    env.current_step = int(step)


def read_state(env):
    world = env.world
    pos = np.stack([a.state.p_pos for a in world.agents]).astype(np.float64)
    vel = np.stack([a.state.p_vel for a in world.agents]).astype(np.float64)
    lm = np.stack([l.state.p_pos for l in world.landmarks]).astype(np.float64)
    return pos, vel, lm


def reference_step(env, act):
    """One ``env.step`` of the unmodified reference.  ``act`` [N,2] is copied first because
    ``_set_action`` scales the caller's array in place (environment.py:216,221)."""
    act_n = [np.array(a, dtype=np.float64) for a in act]
    obs_n, reward_n, done_n, info_n = env.step(act_n)
    pos, vel, lm = read_state(env)
    return dict(
        pos=pos, vel=vel, landmarks=lm,
        obs=np.stack(obs_n).astype(np.float64),
        reward=np.array([r[0] for r in reward_n], dtype=np.float64),
        indiv=np.array([i["individual_reward"] for i in info_n], dtype=np.float64),
        done=np.array(done_n, dtype=bool),
    )


def time_reference(scenario, num_agents, seconds=3.0, episode_length=25, seed=0):
    """Random-policy stepping rate of the unmodified reference (BASELINE.md section 3 loop).
    Returns (env_steps, wall_seconds)."""
    import time
    np.random.seed(seed)
    env = make_reference_env(scenario, num_agents, episode_length)
    env.reset()
    n = num_agents

    def one():
        act_n = [np.random.uniform(-1, 1, 2) for _ in range(n)]
        _, _, done_n, _ = env.step(act_n)
        if np.all(done_n):
            env.reset()

    for _ in range(min(episode_length, 3 if n > 100 else episode_length)):
        one()
    steps, t0 = 0, time.perf_counter()
    while True:
        one()
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return steps, dt


def _timed_worker(q, scenario, num_agents, episode_length, seed, max_steps, seconds, warmup):
    """One env of the reference, random policy U(-1,1) (test.py:20), reset when all agents are done
    (test.py:26-27).  Stops after `max_steps` env steps or `seconds`, whichever comes first."""
    import time
    np.random.seed(seed)
    env = make_reference_env(scenario, num_agents, episode_length)
    env.reset()
    n = num_agents

    def one():
        act_n = [np.random.uniform(-1, 1, 2) for _ in range(n)]
        _, _, done_n, _ = env.step(act_n)
        if np.all(done_n):
            env.reset()

    for _ in range(warmup):
        one()
    steps, t0 = 0, time.perf_counter()
    while steps < max_steps:
        one()
        steps += 1
        if seconds is not None and time.perf_counter() - t0 >= seconds:
            break
    q.put((steps, time.perf_counter() - t0))


def time_reference_parallel(scenario, num_agents, procs=None, seconds=3.0, max_steps=10 ** 9, episode_length=25,
                            warmup=None):
    """The reference's own way of running many envs (one OS process per env, SubprocVecEnv,
    train/maddpg-v2/utils/env_wrappers.py:48-55): `procs` forked processes, one unmodified env each.
    Returns a dict with the aggregate agent-steps/s (sum of env steps x N / slowest worker's wall time)."""
    import multiprocessing as mp
    procs = procs or (os.cpu_count() or 1)
    if warmup is None:
        warmup = 3 if num_agents > 100 else episode_length
    load_reference()                                        # import once in the parent; the forks inherit it
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_timed_worker, args=(q, scenario, num_agents, episode_length, 100 + k, max_steps,
                                                  seconds, warmup)) for k in range(procs)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    env_steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    import scipy
    return {"scenario": scenario, "agents": num_agents, "procs": procs, "env_steps": env_steps, "wall_s": wall,
            "steps_per_worker": [r[0] for r in res][:4],
            "env_steps_per_s": env_steps / wall, "agent_steps_per_s": env_steps * num_agents / wall,
            "reference_root": REFERENCE_ROOT, "numpy": np.__version__, "scipy": scipy.__version__}


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser(description="time the unmodified reference on the host cores (prints one JSON line)")
    ap.add_argument("--scenario", default="formation_hd_env")
    ap.add_argument("--agents", type=int, default=9)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--procs", type=int, default=0, help="0 = os.cpu_count()")
    ap.add_argument("--max-steps", type=int, default=10 ** 9)
    ap.add_argument("--episode-length", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=-1)
    a = ap.parse_args()
    if not reference_available():
        print(json.dumps({"unavailable": "no reference tree at %s" % REFERENCE_ROOT}))
        sys.exit(0)
    r = time_reference_parallel(a.scenario, a.agents, a.procs or None, a.seconds, a.max_steps, a.episode_length,
                                None if a.warmup < 0 else a.warmup)
    print(json.dumps(r))
