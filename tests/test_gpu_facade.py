"""Drop-in API tests on the GPU: ``formation_gym.make_env(...).reset()/step()`` (the reference's
public API, formation_gym/__init__.py:6-17 + environment.py:113-156) driven like test.py drives it,
checked against trajectories frozen from the UNMODIFIED reference (tests/golden/*_traj25.npz) and
against the API contract (types, lengths, aliasing) in tests/golden/api_contract.json.

The facade computes in fp64 (like the reference), so the 25-step bar is 1e-9.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import formation_gym  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def inject(env, g, scenario):
    for a, p, v in zip(env.world.agents, g["pos0"], g["vel0"]):
        a.state.p_pos, a.state.p_vel = p.copy(), v.copy()
        a.state.c = np.zeros(env.world.dim_c)
    sc = env.observation_callback.__self__
    if scenario == "formation_hd_env":
        sc.ideal_shape, sc.ideal_vel = g["shape"].copy(), g["ivel"].copy()
    else:
        for l, p in zip(env.world.landmarks, g["lm0"]):
            l.state.p_pos = p.copy()
    env.current_step = 0


@pytest.mark.parametrize("scenario,name,n", [("formation_hd_env", "hd_n9_traj25.npz", 9),
                                             ("formation_hd_env", "hd_n27_traj25.npz", 27),
                                             ("basic_formation_env", "basic_n3_traj25.npz", 3)])
def test_make_env_step_matches_reference_trajectory(scenario, name, n):
    g = load(name)
    want = json.load(open(os.path.join(GOLD, "api_contract.json")))["api"][scenario]
    np.random.seed(3)
    env = formation_gym.make_env(scenario, False, n, 100)
    obs_n = env.reset()
    assert isinstance(obs_n, list) and len(obs_n) == n and obs_n[0].shape == tuple(env.observation_space[0].shape)
    inject(env, g, scenario)
    worst = 0.0
    for t in range(g["acts"].shape[0]):
        acts = [g["acts"][t][i].copy() for i in range(n)]
        keep = [a.copy() for a in acts]
        obs_n, reward_n, done_n, info_n = env.step(acts)
        # types and structure (environment.py:126-142)
        assert isinstance(obs_n, list) and len(obs_n) == n
        assert obs_n[0].dtype == np.dtype(want["obs_dtype"]) == np.float64
        assert obs_n[0].shape == tuple(env.observation_space[0].shape)
        assert isinstance(reward_n, list) and isinstance(reward_n[0], list) and len(reward_n[0]) == 1
        assert reward_n[0] is reward_n[-1]                       # [[R]] * N aliases one inner list
        assert all(type(d) is bool for d in done_n) and not any(done_n)
        assert sorted(info_n[0].keys()) == want["info_keys"]
        assert all(np.array_equal(a, k) for a, k in zip(acts, keep))   # caller's actions untouched (Q4)
        P = np.stack([a.state.p_pos for a in env.world.agents])
        V = np.stack([a.state.p_vel for a in env.world.agents])
        ind = np.array([i["individual_reward"] for i in info_n])
        worst = max(worst, np.abs(P - g["pos"][t]).max(), np.abs(V - g["vel"][t]).max(),
                    np.abs(ind - g["indiv"][t]).max())
        assert abs(reward_n[0][0] - g["reward"][t][0]) <= 1e-9 * max(1.0, abs(g["reward"][t][0]))
    O = np.stack(obs_n)
    rows = g["obs_rows"] if "obs_rows" in g else np.arange(n)
    worst = max(worst, np.abs(O[rows] - g["obs_last"]).max())
    assert env.current_step == g["acts"].shape[0]
    print("%s N=%d facade 25-step max abs err %.3e" % (scenario, n, worst))
    assert worst <= 1e-9


def test_done_and_reset_like_test_py():
    """test.py:14-28 loop: random policy, reset when all done; done flips at world_length."""
    np.random.seed(5)
    env = formation_gym.make_env("formation_hd_env", False, 9, 4)
    env.seed(5)
    env.reset()
    for t in range(1, 10):
        act_n = [space.sample() for space in env.action_space]
        obs_n, reward_n, done_n, info_n = env.step(act_n)
        assert all(d == (env.current_step >= 4) for d in done_n)
        if np.all(done_n):
            obs_n = env.reset()
            assert env.current_step == 0
            assert all(np.all(a.state.p_vel == 0) for a in env.world.agents)


def test_scenario_hooks_serve_kernel_results():
    """The per-agent hooks (Scenario.observation / reward) evaluate on the GPU and equal what the
    fused env.step reports for the same state; the hd landmark shift side effect is reproduced
    (formation_hd_env.py:40-44)."""
    np.random.seed(9)
    env = formation_gym.make_env("formation_hd_env", False, 9, 25)
    env.reset()
    sc = env.observation_callback.__self__
    obs_n, reward_n, done_n, info_n = env.step([s.sample() for s in env.action_space])
    for i, ag in enumerate(env.world.agents):
        assert np.allclose(sc.observation(ag, env.world), obs_n[i], atol=1e-12)
        assert abs(sc.reward(ag, env.world) - info_n[i]["individual_reward"]) <= 1e-12
    P = np.stack([a.state.p_pos for a in env.world.agents])
    L = np.stack([l.state.p_pos for l in env.world.landmarks])
    assert np.abs(P.mean(0) - L.mean(0)).max() <= 1e-12
    assert np.abs((L - L.mean(0)) - sc.ideal_shape).max() <= 1e-9


def test_custom_callbacks_use_gpu_world_step():
    """User-supplied reward/observation hooks (the plugin API): physics still runs in the kernel
    (World.step), the hooks are called like the reference does (environment.py:126-134)."""
    from formation_gym.environment import MultiAgentEnv
    from formation_gym.envs.formation_hd_env import Scenario
    np.random.seed(1)
    sc = Scenario()
    world = sc.make_world(5, 10)
    calls = {"r": 0}

    def reward(agent, w):
        calls["r"] += 1
        return -float(np.sum(np.square(agent.state.p_pos)))

    def observation(agent, w):
        return np.concatenate([agent.state.p_vel, agent.state.p_pos])

    env = MultiAgentEnv(world, sc.reset_world, reward, observation)
    env.reset()
    p0 = np.stack([a.state.p_pos for a in world.agents]); v0 = np.stack([a.state.p_vel for a in world.agents])
    acts = [np.array([0.3, -0.2]) for _ in range(5)]
    obs_n, reward_n, done_n, info_n = env.step(acts)
    assert calls["r"] == 5 and obs_n[0].shape == (4,)
    # free flight (agents far apart with overwhelming probability): v' = v*0.75 + 5*a*0.1; p' = p + v'*0.1
    v1 = v0 * 0.75 + np.array([1.5, -1.0]) * 0.1
    far = np.min(np.linalg.norm(p0[:, None] - p0[None] + np.eye(5)[..., None] * 9, axis=-1)) > 0.2
    if far:
        assert np.abs(np.stack([a.state.p_vel for a in world.agents]) - v1).max() <= 1e-12
        assert np.abs(np.stack([a.state.p_pos for a in world.agents]) - (p0 + v1 * 0.1)).max() <= 1e-12
    assert abs(reward_n[0][0] - sum(i["individual_reward"] for i in info_n)) <= 1e-12


@pytest.mark.parametrize("scenario", ["formation_hd_env", "basic_formation_env"])
@pytest.mark.parametrize("mode", ["onehot", "input", "force", "box"])
def test_discrete_action_modes_and_benchmark_data(scenario, mode):
    """The less-travelled corners of MultiAgentEnv against the unmodified reference (tests/golden/misc_api.npz,
    make_golden_misc.py): one-hot Discrete(5) action spaces (environment.py:203-206), discrete_action_input
    (:194-202), world.discrete_action -> argmax one-hot (:207-211), and Scenario.benchmark_data."""
    import formation_gym
    from formation_gym.environment import MultiAgentEnv
    g = np.load(os.path.join(GOLD, "misc_api.npz"))
    key = lambda k: g["%s/%s/%s" % (scenario, mode, k)]  # noqa: E731
    n = 3
    env0 = formation_gym.make_env(scenario, False, n)
    sc = env0.reset_callback.__self__
    world = env0.world
    kw = {}
    if mode == "onehot":
        kw = dict(discrete_action=True)
    elif mode == "force":
        world.discrete_action = True
    env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation, **kw)
    if mode == "input":
        env.discrete_action_input = True
    assert int(getattr(env.action_space[0], "n", -1)) == int(key("action_space_n"))
    np.random.seed(77)
    env.reset()
    assert np.array_equal(np.stack([a.state.p_pos for a in world.agents]), key("pos0"))
    for t in range(4):
        a = key("act")[t]
        act_n = [int(x) for x in a] if mode == "input" else [np.array(x, np.float64) for x in a]
        obs_n, reward_n, done_n, info_n = env.step(act_n)
        assert np.abs(np.stack(obs_n) - key("obs")[t]).max() <= 1e-9
        assert abs(reward_n[0][0] - key("reward")[t]) <= 1e-9
        assert np.abs(np.array([i["individual_reward"] for i in info_n]) - key("indiv")[t]).max() <= 1e-9
        assert np.abs(np.stack([a_.state.p_pos for a_ in world.agents]) - key("pos")[t]).max() <= 1e-9
    bench = [sc.benchmark_data(a_, world) for a_ in world.agents]
    got = np.array([[b["reward"], b["collisions"], b["min_dists"], b["occupied_landmarks"]] for b in bench], np.float64)
    assert np.abs(got - key("bench")).max() <= 1e-9


def test_scripted_agent_and_callbacks():
    """A scripted agent (core.py:210-211), done_callback, info_callback's 'fail' key and post_step_callback
    (environment.py:131-133,140-141,172-178) against the unmodified reference: only the policy agents are observed and
    rewarded, the scripted one still pushes and is pushed (World.step on the GPU)."""
    import formation_gym
    from formation_gym.environment import MultiAgentEnv
    from formation_gym.core import Action
    g = np.load(os.path.join(GOLD, "misc_api.npz"))
    key = lambda k: g["scripted/" + k]  # noqa: E731
    env0 = formation_gym.make_env("formation_hd_env", False, 3)
    sc = env0.reset_callback.__self__
    world = env0.world

    def script(agent, w):
        act = Action()
        act.u = np.array([0.3, -0.2]) * (1 + w.world_step)
        act.c = np.zeros(w.dim_c)
        return act
    world.agents[2].action_callback = script
    calls = {"post": 0}
    env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation,
                        info_callback=lambda agent, w: {"fail": agent.state.p_pos[0] > 0.0, "other": 1},
                        done_callback=lambda agent, w: bool(agent.state.p_pos[1] > 0.0),
                        post_step_callback=lambda w: calls.__setitem__("post", calls["post"] + 1))
    assert env.num_agents == int(key("n_policy")) == 2
    np.random.seed(31)
    obs0 = np.stack(env.reset())
    assert np.abs(obs0 - key("obs0")).max() <= 1e-12
    for t in range(4):
        obs_n, reward_n, done_n, info_n = env.step([np.array(x) for x in key("act")[t]])
        assert len(obs_n) == 2 and np.abs(np.stack(obs_n) - key("obs")[t]).max() <= 1e-9
        assert abs(reward_n[0][0] - key("reward")[t]) <= 1e-9
        assert list(done_n) == [bool(x) for x in key("done")[t]]
        assert [bool(i["fail"]) for i in info_n] == [bool(x) for x in key("fail")[t]]
        assert all(set(i.keys()) == {"individual_reward", "fail"} for i in info_n)
        assert np.abs(np.stack([a.state.p_pos for a in world.agents]) - key("pos")[t]).max() <= 1e-9
    assert calls["post"] == int(key("post_calls")) == 4


@pytest.mark.parametrize("scenario,n", [("formation_hd_env", 9), ("basic_formation_env", 3),
                                        ("formation_hd_partial_env", 5), ("formation_hd_partial_range_env", 4),
                                        ("formation_hd_obs_env", 4)])
def test_every_stock_scenario_steps_in_one_fused_launch(scenario, n):
    """env.step of every stock scenario = ONE kernel launch (fg_step_fused), and the host records of the actions
    are what the reference's _set_action leaves (environment.py:188-236: u = action * sensitivity, c = 0)."""
    np.random.seed(2)
    env = formation_gym.make_env(scenario, False, n, 10)
    env.reset()
    be = env.world.backend()
    for _ in range(3):
        before = be.launches
        acts = [np.random.uniform(-1, 1, 2) for _ in range(n)]
        obs_n, reward_n, done_n, info_n = env.step([a.copy() for a in acts])
        assert be.launches == before + 1
        for a, ag in zip(acts, env.world.agents):
            assert np.allclose(ag.action.u, a * 5.0, rtol=0, atol=0) and np.all(ag.action.c == 0)
    assert len(obs_n) == n and obs_n[0].shape == tuple(env.observation_space[0].shape)


def test_hook_cache_sees_changed_constants():
    """The scenario hooks share one cached launch per state; the cache key must cover the constants the kernels read
    (round-1 advice: agent.size sets the reward's collision threshold)."""
    np.random.seed(4)
    env = formation_gym.make_env("formation_hd_env", False, 3, 10)
    env.reset()
    w, sc = env.world, env.observation_callback.__self__
    w.agents[0].state.p_pos = np.array([0.0, 0.0]); w.agents[1].state.p_pos = np.array([0.04, 0.0])
    r_small = sc.reward(w.agents[0], w)                  # |dp| = 0.04 >= (0.03 + 0.03) / 2: no collision
    for a in w.agents:
        a.size = 0.05                                    # threshold (0.05 + 0.05) / 2 = 0.05 > 0.04: collision
    r_big = sc.reward(w.agents[0], w)
    assert abs((r_small - r_big) - 1.0) <= 1e-12


def test_cache_dists_matches_reference():
    """World.cache_dists (core.py:132,156-180,224-225,298-301): calculate_distances() runs on the device and fills the
    reference's four cache arrays; with the cache on, a step's contact forces come from the positions of the previous
    calculate_distances() -- visible when the state is edited in between (agent 1 is teleported next to agent 0: the
    next step misses the contact, the one after sees it).  Fixture from the unmodified reference, fp64, 1e-12."""
    g = load("cache_dists_n5.npz")
    np.random.seed(1)
    env = formation_gym.make_env("formation_hd_env", False, 5, 25)
    env.reset()
    inject(env, g, "formation_hd_env")
    w = env.world
    for l, p in zip(w.landmarks, g["lm0"]):
        l.state.p_pos = p.copy()
    w.cache_dists = True
    with pytest.raises(TypeError):                       # the reference subscripts the empty cache (core.py:299)
        env.step([np.zeros(2)] * 5)
    env.current_step = 0
    w.calculate_distances()
    assert np.abs(w.cached_dist_vect - g["vect0"]).max() <= 1e-15 and np.abs(w.cached_dist_mag - g["mag0"]).max() <= 1e-15
    assert np.array_equal(w.min_dists, g["mind"]) and np.array_equal(w.cached_collisions, g["coll0"])
    assert w.cached_collisions.dtype == np.bool_ and w.cached_collisions.diagonal().all()
    for t in range(g["acts"].shape[0]):
        if t == 3:
            w.agents[1].state.p_pos = w.agents[0].state.p_pos + np.array([0.04, 0.01])
            assert np.abs(np.stack([a.state.p_pos for a in w.agents]) - g["tele_pos"]).max() <= 1e-12
        obs_n, reward_n, done_n, info_n = env.step([a.copy() for a in g["acts"][t]])
        P = np.stack([a.state.p_pos for a in w.agents]); V = np.stack([a.state.p_vel for a in w.agents])
        assert np.abs(P - g["pos"][t]).max() <= 1e-12 and np.abs(V - g["vel"][t]).max() <= 1e-12, t
        assert abs(reward_n[0][0] - g["reward"][t]) <= 1e-11 and np.abs(np.stack(obs_n) - g["obs"][t]).max() <= 1e-12
    assert np.abs(w.cached_dist_vect - g["vect_last"]).max() <= 1e-12
    assert np.abs(w.cached_dist_mag - g["mag_last"]).max() <= 1e-12 and np.array_equal(w.cached_collisions, g["coll_last"])
