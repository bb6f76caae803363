"""Golden fixtures for the less-travelled corners of the reference API (build container only; unmodified reference):
discrete action spaces of MultiAgentEnv (environment.py:194-210), discrete_action_input, world.discrete_action
(force_discrete_action), and Scenario.benchmark_data (formation_hd_env.py:97-117, basic_formation_env.py:67-87).
    python tests/golden/make_golden_misc.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402


def build(scenario, n, **env_kw):
    fg = rh.load_reference()
    from formation_gym.environment import MultiAgentEnv          # the reference's class
    env0 = fg.make_env(scenario, False, n)
    sc = rh.scenario_of(env0)
    world = env0.world
    for k, v in env_kw.pop("world_attrs", {}).items():
        setattr(world, k, v)
    env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation, **env_kw)
    return env, sc


def run(scenario, n, mode, seed):
    rng = np.random.default_rng(seed)
    kw = {}
    if mode == "onehot":
        kw = dict(discrete_action=True)
    elif mode == "force":
        kw = dict(world_attrs={"discrete_action": True})
    env, sc = build(scenario, n, **kw)
    if mode == "input":
        env.discrete_action_input = True
    np.random.seed(seed)
    env.reset()
    pos0 = np.stack([a.state.p_pos for a in env.world.agents])
    lm0 = np.stack([l.state.p_pos for l in env.world.landmarks])
    extra = {}
    if scenario == "formation_hd_env":
        extra = dict(shape=np.array(sc.ideal_shape), ivel=np.array(sc.ideal_vel))
    acts, outs = [], dict(obs=[], reward=[], indiv=[], pos=[])
    for t in range(4):
        if mode == "onehot":
            a = np.zeros((n, 5)); a[np.arange(n), rng.integers(0, 5, n)] = 1.0
            a += 0.1 * rng.uniform(0, 1, (n, 5))                     # soft one-hot: u = [a1 - a2, a3 - a4]
        elif mode == "input":
            a = rng.integers(0, 5, n).astype(np.int64)
        else:
            a = rng.uniform(-1, 1, (n, 2))
        acts.append(np.array(a, np.float64))
        act_n = [int(x) for x in a] if mode == "input" else [np.array(x, np.float64) for x in a]
        obs_n, reward_n, done_n, info_n = env.step(act_n)
        outs["obs"].append(np.stack(obs_n)); outs["reward"].append(reward_n[0][0])
        outs["indiv"].append([i["individual_reward"] for i in info_n])
        outs["pos"].append(np.stack([a_.state.p_pos for a_ in env.world.agents]))
    bench = [sc.benchmark_data(a_, env.world) for a_ in env.world.agents]
    d = dict(pos0=pos0, lm0=lm0, act=np.stack(acts), **{k: np.array(v, np.float64) for k, v in outs.items()}, **extra)
    d["bench"] = np.array([[b["reward"], b["collisions"], b["min_dists"], b["occupied_landmarks"]] for b in bench], np.float64)
    d["action_space_n"] = np.int64(getattr(env.action_space[0], "n", -1))
    return d


def run_scripted(seed):
    """formation_hd_env, 3 agents, agent 2 scripted (core.py:210-211: action_callback), callbacks that count calls:
    done_callback, post_step_callback, info_callback with a 'fail' key (environment.py:131-133,140-141)."""
    fg = rh.load_reference()
    from formation_gym.environment import MultiAgentEnv
    from formation_gym.core import Action
    env0 = fg.make_env("formation_hd_env", False, 3)
    sc = rh.scenario_of(env0)
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
    np.random.seed(seed)
    obs0 = np.stack(env.reset())
    rng = np.random.default_rng(seed)
    d = dict(obs0=obs0, n_policy=np.int64(env.num_agents), obs=[], reward=[], done=[], fail=[], pos=[], act=[])
    for t in range(4):
        a = rng.uniform(-1, 1, (2, 2))
        obs_n, reward_n, done_n, info_n = env.step([np.array(x) for x in a])
        d["act"].append(a); d["obs"].append(np.stack(obs_n)); d["reward"].append(reward_n[0][0])
        d["done"].append(done_n); d["fail"].append([bool(i["fail"]) for i in info_n])
        d["pos"].append(np.stack([ag.state.p_pos for ag in world.agents]))
        assert all(set(i.keys()) == {"individual_reward", "fail"} for i in info_n)
    d["post_calls"] = np.int64(calls["post"])
    return {k: np.array(v) for k, v in d.items()}


def run_bfs_inconsistent(seed, n_layer=3, layers=2):
    """get_action_BFS(ezpolicy, obs_n, 3) of the unmodified reference (formation_gym/__init__.py:19-98) on
    observations that do NOT describe one consistent state: some agents' other_pos slices are perturbed (what a
    noisy / clipping observation wrapper produces).  The reference reads every leader's and member's OWN
    observation, so the result differs from what agent 0's observation alone would give."""
    fg = rh.load_reference()
    N = n_layer ** layers
    np.random.seed(seed)
    env = fg.make_env("formation_hd_env", False, N)
    obs_n = [np.array(o, np.float64) for o in env.reset()]
    rng = np.random.default_rng(seed)
    clean = np.stack(obs_n)
    act_clean = np.stack(fg.get_action_BFS(fg.ezpolicy, [o.copy() for o in obs_n], n_layer))
    for j in (1, 3, 4, 8):
        obs_n[j][2:2 * N] += rng.normal(0.0, 0.2, 2 * N - 2)
    noisy = np.stack(obs_n)
    act_noisy = np.stack(fg.get_action_BFS(fg.ezpolicy, [o.copy() for o in obs_n], n_layer))
    assert np.abs(act_noisy - act_clean).max() > 1e-3
    return dict(obs_clean=clean, act_clean=act_clean, obs_noisy=noisy, act_noisy=act_noisy, n=np.int64(n_layer))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "bfs_inconsistent_n9.npz"), **run_bfs_inconsistent(5))
    out = {}
    for k, v in run_scripted(31).items():
        out["scripted/" + k] = v
    print("scripted", out["scripted/obs"].shape, int(out["scripted/n_policy"]), int(out["scripted/post_calls"]))
    for scenario, n in (("formation_hd_env", 3), ("basic_formation_env", 3)):
        for mode in ("onehot", "input", "force", "box"):
            d = run(scenario, n, mode, 77)
            for k, v in d.items():
                out["%s/%s/%s" % (scenario, mode, k)] = v
            print(scenario, mode, d["obs"].shape, d["bench"][0])
    np.savez_compressed(os.path.join(HERE, "misc_api.npz"), **out)
