"""formation_hd_obs_env (movable colliding obstacle landmarks) on the CUDA path, against
fixtures frozen from the unmodified reference (tests/golden/obstacle_*.npz) and the numpy oracle."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import formation_gym  # noqa: E402
from formation_gym.batched import BatchedFormationEnv  # noqa: E402
from oracle import mpe_oracle as mo  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "obstacle_n*.npz")))
SCN = "formation_hd_obs_env"


def _dev(dtype):
    return lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")


def _err(a, b):
    return float(np.abs(a.double().cpu().numpy() - b).max())


def _load(env, dev, pos, vel, goals, obst, obst_vel, step0):
    G = env.num_goals
    env.pos.copy_(dev(pos)); env.vel.copy_(dev(vel))
    env.landmarks[:, :G].copy_(dev(goals)); env.landmarks[:, G:].copy_(dev(obst))
    env.landmark_vel.zero_(); env.landmark_vel[:, G:].copy_(dev(obst_vel))
    env.step_count.copy_(torch.as_tensor(step0, dtype=torch.int32, device="cuda"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_obstacle_single_step_golden(path, dtype):
    g = np.load(path)
    E, N = g["pos0"].shape[:2]
    G, O = g["goals"].shape[1], g["obst0"].shape[1]
    env = BatchedFormationEnv(SCN, E, N, episode_length=50, num_landmarks=G, num_obstacles=O, dtype=dtype,
                              auto_reset=False)
    dev = _dev(dtype)
    _load(env, dev, g["pos0"], g["vel0"], g["goals"], g["obst0"], g["obst_vel0"], g["step0"])
    obs, rew, done, info = env.step(dev(g["act"]))
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert obs.shape[2] == g["obs"].shape[2]
    assert _err(env.pos, g["pos"]) <= tol and _err(env.vel, g["vel"]) <= tol
    assert _err(env.landmarks[:, G:], g["obst"]) <= tol and _err(env.landmarks[:, :G], g["goals"]) == 0.0
    assert _err(env.landmark_vel[:, G:], g["obst_vel"]) == 0.0          # the reward hook's rule: exact values
    assert _err(obs, g["obs"]) <= tol
    assert _err(info["individual_reward"], g["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
    assert _err(rew[:, :, 0], g["reward"]) <= tol * (8 if dtype == torch.float32 else 100)
    assert np.array_equal(done.cpu().numpy(), g["done"])
    # observation + reward alone (fg_obs_reward) from the stepped state gives the same rows
    obs2 = env.observe().clone()
    assert torch.equal(obs2, obs)


def test_obstacle_trajectory_fp64():
    """50 steps from the reference's seeded reset; the obstacles fall through the agents (contacts)."""
    g = np.load(os.path.join(GOLD, "obstacle_traj50.npz"))
    T, N = g["act"].shape[:2]
    G, O = g["goals"].shape[0], g["obst0"].shape[0]
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-4)):
        env = BatchedFormationEnv(SCN, 1, N, episode_length=50, num_landmarks=G, num_obstacles=O, dtype=dtype,
                                  auto_reset=False)
        dev = _dev(dtype)
        _load(env, dev, g["pos0"][None], np.zeros((1, N, 2)), g["goals"][None], g["obst0"][None],
              g["obst_vel0"][None], np.zeros(1))
        assert _err(env.observe()[0], g["obs0"]) <= (1e-12 if dtype == torch.float64 else 1e-6)
        worst = 0.0
        for t in range(T):
            obs, rew, done, info = env.step(dev(g["act"][t][None]))
            worst = max(worst, _err(env.pos[0], g["pos"][t]), _err(env.landmarks[0, G:], g["obst"][t]),
                        _err(obs[0], g["obs"][t]))
            if dtype == torch.float64:
                assert _err(info["individual_reward"][0], g["indiv"][t]) <= 1e-9
            assert _err(env.landmark_vel[0, G:], g["obst_vel"][t]) == 0.0
        assert worst <= tol, worst
        assert bool(done.all())


@pytest.mark.parametrize("E,N,G,O", [(300, 4, 4, 3), (33, 40, 7, 5), (5, 243, 9, 4), (64, 9, 1, 1), (7, 256, 256 - 8, 8)])
def test_obstacle_vs_oracle_random(E, N, G, O):
    rng = np.random.default_rng(N * 17 + O)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-0.7, 0.7, (E, N, 2))); vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2))); goals = f32(rng.uniform(-1, 1, (E, G, 2)))
    obst = f32(rng.uniform(-0.9, 0.9, (E, O, 2))); ov = f32(rng.uniform(-1, 1, (E, O, 2)))
    obst[::3, :, 1] -= 2.6
    step0 = rng.integers(0, 50, E)
    ref = mo.obstacle_env_step(pos, vel, act, goals, obst, ov, step0)
    for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 1e-11)):
        env = BatchedFormationEnv(SCN, E, N, episode_length=50, num_landmarks=G, num_obstacles=O, dtype=dtype,
                                  auto_reset=False)
        dev = _dev(dtype)
        _load(env, dev, pos, vel, goals, obst, ov, step0)
        obs, rew, done, info = env.step(dev(act))
        assert _err(env.pos, ref["pos"]) <= tol and _err(env.vel, ref["vel"]) <= tol * 10
        assert _err(env.landmarks[:, G:], ref["obst"]) <= tol
        assert _err(env.landmark_vel[:, G:], ref["obst_vel"]) == 0.0
        assert _err(obs, ref["obs"]) <= tol * 10
        assert _err(info["individual_reward"], ref["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
        R = ref["reward"]
        assert np.all(np.abs(rew[:, 0, 0].double().cpu().numpy() - R) <= tol + 1e-6 * np.abs(R))
        assert np.array_equal(done[:, 0].cpu().numpy(), ref["done"])


def test_obstacle_reset_rollout_and_facade():
    """Device reset distribution (formation_hd_obs_env.py:101-120), in-kernel rollout == stepwise, auto-reset,
    and the single-env facade through the reference API against the frozen reference trajectory."""
    env = BatchedFormationEnv(SCN, 4096, 4, episode_length=5, seed=3)
    env.reset()
    G, O = env.num_goals, env.num_obstacles
    ob = env.landmarks[:, G:].cpu().numpy()
    edges = np.linspace(-1.8, 1.8, O + 1)
    for k in range(O):
        assert ob[:, k, 0].min() >= edges[k] and ob[:, k, 0].max() <= edges[k + 1]
        assert abs(ob[:, k, 0].mean() - 0.5 * (edges[k] + edges[k + 1])) < 0.03
    assert ob[..., 1].min() >= 2.0 and ob[..., 1].max() <= 2.5
    ov = env.landmark_vel[:, G:].cpu().numpy()
    assert np.all(ov[..., 0] == 0.0) and np.all(ov[..., 1] == -1.0)
    assert float(env.landmarks[:, :G].abs().max()) <= 1.0 and float(env.pos.abs().max()) <= 1.0
    # stepwise random policy == one in-kernel rollout (same Philox stream), across an auto-reset
    # (fp32: both through the tile kernel -- single steps would otherwise take the warp kernel's STD instantiation,
    # whose FMA contraction differs in the last bit)
    from formation_gym import _native as nat_
    sd = env.state_dict()
    with nat_.options(no_std_kernel=1):
        for _ in range(7):
            obs_a, rew_a, done_a, _ = env.step_random()
    fin = {k: getattr(env, k).clone() for k in ("pos", "vel", "landmarks", "landmark_vel", "step_count")}
    obs_a = obs_a.clone()
    env.load_state_dict(sd)
    obs_b, rew_b, done_b, _ = env.rollout_random(7)
    for k, v in fin.items():
        assert torch.equal(getattr(env, k), v), k
    assert torch.equal(obs_a, obs_b)
    assert int(env.step_count[0]) == 2                                  # 7 steps, episodes of 5
    # facade: reference API, trajectory frozen from the unmodified reference
    g = np.load(os.path.join(GOLD, "obstacle_traj50.npz"))
    fenv = formation_gym.make_env(SCN, False, 4)
    assert fenv.world_length == 50 and len(fenv.world.landmarks) == 7
    np.random.seed(7)
    obs_n = fenv.reset()
    assert np.abs(np.stack(obs_n) - g["obs0"]).max() <= 1e-12
    for t in range(6):
        obs_n, reward_n, done_n, info_n = fenv.step([a.copy() for a in g["act"][t]])
        assert np.abs(np.stack(obs_n) - g["obs"][t]).max() <= 1e-9
        assert abs(reward_n[0][0] - g["reward"][t][0]) <= 1e-9 and reward_n[0] is reward_n[1]
        lm = np.stack([l.state.p_pos for l in fenv.world.landmarks])
        assert np.abs(lm[4:] - g["obst"][t]).max() <= 1e-9
        lv = np.stack([l.state.p_vel for l in fenv.world.landmarks[4:]])
        assert np.array_equal(lv, g["obst_vel"][t])
    # the hooks alone (observation / reward from the host state), like user code calling them
    sc = fenv.reward_callback.__self__
    r0 = sc.reward(fenv.world.agents[0], fenv.world)
    assert abs(r0 - g["indiv"][5][0]) <= 1e-9


def test_world_step_alone_on_obstacle_world():
    """World.step() without the scenario hooks (core.py:206-225) on a world with movable colliding landmarks:
    fg_world_step integrates the obstacles too and leaves their INTEGRATED velocity (the (0, -1) rule belongs to the
    reward hook).  Batched entry point and the facade's World.step() against the oracle."""
    rng = np.random.default_rng(5)
    E, N, G, O = 40, 6, 3, 4
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-0.5, 0.5, (E, N, 2))); vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2))); goals = f32(rng.uniform(-1, 1, (E, G, 2)))
    obst = f32(rng.uniform(-0.6, 0.6, (E, O, 2))); ov = f32(rng.uniform(-1, 1, (E, O, 2)))
    p, v, o, ovi = mo.obstacle_world_step(pos, vel, act, obst, ov)
    for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 1e-11)):
        env = BatchedFormationEnv(SCN, E, N, num_landmarks=G, num_obstacles=O, dtype=dtype, auto_reset=False)
        dev = _dev(dtype)
        _load(env, dev, pos, vel, goals, obst, ov, np.zeros(E))
        env.world_step(dev(act))
        assert _err(env.pos, p) <= tol and _err(env.vel, v) <= 10 * tol
        assert _err(env.landmarks[:, G:], o) <= tol and _err(env.landmark_vel[:, G:], ovi) <= 10 * tol
        assert _err(env.landmarks[:, :G], goals) == 0.0
    # facade: World.step() as user code with its own callbacks would call it
    fenv = formation_gym.make_env(SCN, False, 4)
    w = fenv.world
    P0 = np.stack([a.state.p_pos for a in w.agents]) * 0.3
    for a, q in zip(w.agents, P0):
        a.state.p_pos = q.copy()
        a.action.u = np.array([0.5, -0.25]) * 5.0                   # already scaled by _set_action
    lm = np.stack([l.state.p_pos for l in w.landmarks])
    lm[4:] = np.array([[0.05, 0.1], [-0.2, 0.0], [0.3, -0.1]])      # obstacles among the agents
    for l, q in zip(w.landmarks, lm):
        l.state.p_pos = q.copy()
    lv = np.stack([np.asarray(l.state.p_vel, float) for l in w.landmarks[4:]])
    V0 = np.stack([a.state.p_vel for a in w.agents])
    w.step()
    rp, rv, ro, rov = mo.obstacle_world_step(P0[None], V0[None], np.tile([0.5, -0.25], (1, 4, 1)), lm[None, 4:], lv[None])
    assert np.abs(np.stack([a.state.p_pos for a in w.agents]) - rp[0]).max() <= 1e-11
    assert np.abs(np.stack([l.state.p_pos for l in w.landmarks[4:]]) - ro[0]).max() <= 1e-11
    assert np.abs(np.stack([l.state.p_vel for l in w.landmarks[4:]]) - rov[0]).max() <= 1e-10
    assert np.abs(rov[0] - np.array([0.0, -0.75])).max() > 1e-3      # contacts did change an obstacle's velocity


# ---------------------------------------------------------------------------------------------------------------
# k_lm_warp<.., kScnObstacle> (fg_warp_lm.cuh): the warp-autonomous kernel the scenario takes at 3 .. 9 agents with
# make_world's 4 goals + 3 obstacles, against the tile kernel of fg_obstacle.cuh (force_tile_kernel)
from formation_gym import _native as nat  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8, 9])
def test_obstacle_warp_kernel_equals_tile_kernel(N, dtype):
    """Episodes of 12 steps with auto-reset (the obstacles fall through the agents, which start inside [-1, 1]^2: contacts
    agent-agent, agent-obstacle and obstacle-obstacle all occur), ragged last span, random policy recorded: the fp64
    build is bit-identical, fp32 agrees to 1e-5 with the states re-synchronised every step."""
    E = 517
    mk = lambda: BatchedFormationEnv(SCN, E, N, episode_length=12, dtype=dtype, seed=23, auto_reset=True)  # noqa: E731
    a, b = mk(), mk()
    a.reset(); b.reset()
    assert torch.equal(a.obs, b.obs) and torch.equal(a.landmarks, b.landmarks)
    touched = 0
    for t in range(30):
        l0 = a.launches
        oa, ra, da, ia = a.step_random(record_actions=True)
        assert a.launches - l0 == 1
        with nat.options(force_tile_kernel=1):
            ob, rb, db, ib = b.step_random(record_actions=True)
        assert torch.equal(a.actions, b.actions) and torch.equal(da, db)
        pairs = ((oa, ob), (ra, rb), (a.pos, b.pos), (a.vel, b.vel), (a.landmarks, b.landmarks),
                 (a.landmark_vel, b.landmark_vel), (ia["individual_reward"], ib["individual_reward"]),
                 (a.ep_return, b.ep_return))
        if dtype == torch.float64:
            for k, (x, y) in enumerate(pairs):
                assert torch.equal(x, y), (t, k)
        else:
            for k, (x, y) in enumerate(pairs):
                assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max())), (t, k)
            b.load_state_dict(a.state_dict())
        assert torch.equal(a.step_count, b.step_count) and torch.equal(a.ep_collisions, b.ep_collisions)
        touched += int((ia["individual_reward"] < ia["individual_reward"].amax(dim=1, keepdim=True) - 1.5).sum())
    assert touched > 0                                               # some agent paid the -2 of a collision
    assert torch.equal(a.stats[:1], b.stats[:1])


def test_obstacle_warp_kernel_rollout_and_external_actions():
    """n_steps > 1 in one launch (the obstacles' new positions feed the next step's contacts) and caller-owned action
    buffers, fp64, against single tile-kernel steps."""
    for N in (4, 5):
        mk = lambda: BatchedFormationEnv(SCN, 700, N, episode_length=9, dtype=torch.float64, seed=4)  # noqa: E731
        a, b = mk(), mk()
        a.reset(); b.reset()
        a.rollout_random(14)
        for _ in range(14):
            with nat.options(force_tile_kernel=1):
                b.step_random()
        for k in ("pos", "vel", "landmarks", "landmark_vel", "obs", "reward", "step_count", "ep_return"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (N, k)
        act = torch.rand(700, N, 2, dtype=torch.float64, device="cuda") * 2 - 1
        oa = a.step(act)[0].clone()
        with nat.options(force_tile_kernel=1):
            ob = b.step(act)[0]
        assert torch.equal(oa, ob) and torch.equal(a.landmarks, b.landmarks)
