"""formation_hd_partial_env / formation_hd_partial_range_env on the CUDA path, against
fixtures frozen from the unmodified reference (tests/golden/{partial,range}_n*.npz) and the numpy oracle."""
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
FILES = sorted(glob.glob(os.path.join(GOLD, "partial_n*.npz")) + glob.glob(os.path.join(GOLD, "range_n*.npz")))


def _env(name, g, dtype, E=None, N=None, L=None, **kw):
    scen = "formation_hd_partial_env" if name.startswith("partial") else "formation_hd_partial_range_env"
    extra = dict(num_obs=int(g["num_obs"])) if name.startswith("partial") else dict(obs_range=float(g["obs_range"]))
    extra.update(kw)
    return BatchedFormationEnv(scen, E, N, episode_length=25, num_landmarks=L, dtype=dtype, auto_reset=False, **extra)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_partial_single_step_golden(path, dtype):
    g = np.load(path)
    name = os.path.basename(path)
    E, N = g["pos0"].shape[:2]
    L = g["lm"].shape[1]
    env = _env(name, g, dtype, E, N, L)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")  # noqa: E731
    env.pos.copy_(dev(g["pos0"])); env.vel.copy_(dev(g["vel0"])); env.landmarks.copy_(dev(g["lm"]))
    env.step_count.copy_(torch.as_tensor(g["step0"], dtype=torch.int32, device="cuda"))
    obs, rew, done, info = env.step(dev(g["act"]))
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    err = lambda a, b: float(np.abs(a.double().cpu().numpy() - b).max())  # noqa: E731
    assert obs.shape[2] == g["obs"].shape[2]
    assert err(env.pos, g["pos"]) <= tol and err(env.vel, g["vel"]) <= tol
    assert err(obs, g["obs"]) <= tol
    assert err(info["individual_reward"], g["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
    assert err(rew[:, :, 0], g["reward"]) <= tol * (4 if dtype == torch.float32 else 100)
    assert np.array_equal(done.cpu().numpy(), g["done"])


@pytest.mark.parametrize("scen,E,N,L,kw", [("formation_hd_partial_env", 300, 9, 5, dict(num_obs=3)),
                                           ("formation_hd_partial_env", 33, 40, 7, dict(num_obs=12)),
                                           ("formation_hd_partial_env", 5, 243, 243, dict(num_obs=2)),
                                           ("formation_hd_partial_range_env", 300, 9, 4, dict(obs_range=0.7)),
                                           ("formation_hd_partial_range_env", 17, 81, 81, dict(obs_range=0.25))])
def test_partial_vs_oracle_random(scen, E, N, L, kw):
    rng = np.random.default_rng(N * 13 + L)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-0.4, 0.4, (E, N, 2))); vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2))); lm = f32(rng.uniform(-1, 1, (E, L, 2)))
    step0 = rng.integers(0, 25, E)
    ref = mo.partial_env_step(pos, vel, act, lm, step0, num_obs=kw.get("num_obs"), obs_range=kw.get("obs_range"))
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-11)):
        env = BatchedFormationEnv(scen, E, N, episode_length=25, num_landmarks=L, dtype=dtype, auto_reset=False, **kw)
        dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")  # noqa: E731
        env.pos.copy_(dev(pos)); env.vel.copy_(dev(vel)); env.landmarks.copy_(dev(lm))
        env.step_count.copy_(torch.as_tensor(step0, dtype=torch.int32, device="cuda"))
        obs, rew, done, info = env.step(dev(act))
        err = lambda a, b: float(np.abs(a.double().cpu().numpy() - b).max())  # noqa: E731
        assert err(env.pos, ref["pos"]) <= tol and err(obs, ref["obs"]) <= tol
        assert err(info["individual_reward"], ref["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
        R = ref["reward"]
        assert np.all(np.abs(rew[:, 0, 0].double().cpu().numpy() - R) <= tol + 1e-6 * np.abs(R))
        assert np.array_equal(done[:, 0].cpu().numpy(), ref["done"])


def test_partial_rollout_auto_reset_and_facade():
    """Batched rollout with auto-reset (landmarks redrawn, step counter reset) and the single-env facade
    through the reference API."""
    env = BatchedFormationEnv("formation_hd_partial_env", 64, 5, episode_length=4, seed=3)
    env.reset()
    lm0 = env.landmarks.clone()
    for t in range(4):
        obs, rew, done, info = env.step_random()
    assert bool(done.all()) and int(env.step_count.abs().sum()) == 0
    assert not torch.equal(lm0, env.landmarks)                      # reset_world drew new landmarks
    assert torch.equal(env.observe(), obs)                          # the returned obs is the reset state's obs
    fenv = formation_gym.make_env("formation_hd_partial_range_env", False, 4, episode_length=6)
    fenv.seed(1)
    obs_n = fenv.reset()
    assert len(obs_n) == 4 and obs_n[0].shape == (2 + 2 * 4 + 4 * 3,)
    obs_n, reward_n, done_n, info_n = fenv.step([np.array([0.1, -0.2])] * 4)
    assert len(reward_n) == 4 and reward_n[0] is reward_n[1] and done_n == [False] * 4
    assert 'individual_reward' in info_n[0]


# ---------------------------------------------------------------------------------------------------------------
# k_lm_warp (fg_warp_lm.cuh): the warp-autonomous kernel these scenarios take at 3 .. 9 agents with make_world's
# default landmark counts, against the generic tile kernel (force_tile_kernel) and the oracle
from formation_gym import _native as nat  # noqa: E402

LM_CASES = [("formation_hd_partial_env", N) for N in range(3, 10)] + \
           [("formation_hd_partial_range_env", N) for N in range(3, 10)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("scen,N", LM_CASES)
def test_lm_warp_kernel_equals_tile_kernel(scen, N, dtype):
    """Rollouts over auto-resets, ragged last span, random policy recorded: fp64 bit-identical, fp32 to 1e-5 (FMA
    contraction differs between the kernels), with the states re-synchronised every step."""
    E = 1003
    mk = lambda: BatchedFormationEnv(scen, E, N, episode_length=5, dtype=dtype, seed=17, auto_reset=True)  # noqa: E731
    a, b = mk(), mk()
    a.reset(); b.reset()
    assert torch.equal(a.obs, b.obs)
    for t in range(12):
        l0 = a.launches
        oa, ra, da, ia = a.step_random(record_actions=True)
        assert a.launches - l0 == 1
        with nat.options(force_tile_kernel=1):
            ob, rb, db, ib = b.step_random(record_actions=True)
        assert torch.equal(a.actions, b.actions) and torch.equal(da, db)
        pairs = ((oa, ob), (ra, rb), (a.pos, b.pos), (a.vel, b.vel), (a.landmarks, b.landmarks),
                 (ia["individual_reward"], ib["individual_reward"]), (a.ep_return, b.ep_return))
        if dtype == torch.float64:
            for k, (x, y) in enumerate(pairs):
                assert torch.equal(x, y), (t, k)
        else:
            for k, (x, y) in enumerate(pairs):
                assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max())), (t, k)
            b.load_state_dict(a.state_dict())
        assert torch.equal(a.step_count, b.step_count) and torch.equal(a.ep_collisions, b.ep_collisions)
    assert torch.equal(a.stats[:1], b.stats[:1])


def test_lm_warp_kernel_is_the_one_that_runs():
    """The default dispatch takes k_lm_warp for the reference's recipe sizes: its launch geometry (persistent warps,
    <= one wave) differs from the tile kernel's one CTA per tile, which shows in the results only through timing, so
    check the dispatch through the rollout entry (n_steps > 1 in ONE launch) and external action buffers instead."""
    for scen, N in (("formation_hd_partial_env", 4), ("formation_hd_partial_env", 5), ("formation_hd_partial_range_env", 4)):
        mk = lambda: BatchedFormationEnv(scen, 700, N, episode_length=6, dtype=torch.float64, seed=2)  # noqa: E731
        a, b = mk(), mk()
        a.reset(); b.reset()
        a.rollout_random(9)
        for _ in range(9):
            with nat.options(force_tile_kernel=1):
                b.step_random()
        for k in ("pos", "vel", "landmarks", "obs", "reward", "step_count", "ep_return"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (scen, N, k)
        act = torch.rand(700, N, 2, dtype=torch.float64, device="cuda") * 2 - 1
        oa = a.step(act)[0].clone()
        with nat.options(force_tile_kernel=1):
            ob = b.step(act)[0]
        assert torch.equal(oa, ob)


@pytest.mark.parametrize("scen,N,L,kw", [("formation_hd_partial_env", 4, 5, dict(num_obs=3)),
                                         ("formation_hd_partial_env", 5, 5, dict(num_obs=3)),
                                         ("formation_hd_partial_env", 3, 5, dict(num_obs=3)),
                                         ("formation_hd_partial_env", 7, 5, dict(num_obs=3)),
                                         ("formation_hd_partial_range_env", 4, 4, dict(obs_range=0.7)),
                                         ("formation_hd_partial_range_env", 6, 4, dict(obs_range=0.3))])
def test_lm_warp_kernel_vs_oracle(scen, N, L, kw):
    """Seeded random batches (dense: contacts and collisions occur) against the numpy oracle; num_obs = 3 >= N wraps
    onto the agent itself at N = 3 (formation_hd_partial_env.py:50: i % num_agents)."""
    E = 410
    rng = np.random.default_rng(N * 31 + L)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-0.25, 0.25, (E, N, 2))); vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2))); lm = f32(rng.uniform(-1, 1, (E, L, 2)))
    step0 = rng.integers(0, 25, E)
    ref = mo.partial_env_step(pos, vel, act, lm, step0, num_obs=kw.get("num_obs"), obs_range=kw.get("obs_range"))
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-11)):
        env = BatchedFormationEnv(scen, E, N, episode_length=25, num_landmarks=L, dtype=dtype, auto_reset=False, **kw)
        dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")  # noqa: E731
        env.pos.copy_(dev(pos)); env.vel.copy_(dev(vel)); env.landmarks.copy_(dev(lm))
        env.step_count.copy_(torch.as_tensor(step0, dtype=torch.int32, device="cuda"))
        obs, rew, done, info = env.step(dev(act))
        err = lambda a, b: float(np.abs(a.double().cpu().numpy() - b).max())  # noqa: E731
        assert err(env.pos, ref["pos"]) <= tol and err(obs, ref["obs"]) <= tol
        assert err(info["individual_reward"], ref["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
        R = ref["reward"]
        assert np.all(np.abs(rew[:, 0, 0].double().cpu().numpy() - R) <= tol + 1e-6 * np.abs(R))
        assert np.array_equal(done[:, 0].cpu().numpy(), ref["done"])
        assert int((ref["indiv"] < ref["indiv"].max(axis=1, keepdims=True) - 0.5).sum()) > 0   # collisions were exercised


@pytest.mark.parametrize("scen,N,kw", [
    ("formation_hd_partial_env", 4, dict(u_noise=0.1)),
    ("formation_hd_partial_env", 5, dict(max_speed=0.3, mass=2.0)),
    ("formation_hd_partial_range_env", 4, dict(collide=False, accel=3.0)),
    ("formation_hd_obs_env", 4, dict(u_noise=0.05, max_speed=0.6, obstacle_mass=2.5)),
    ("formation_hd_obs_env", 5, dict(mass=0.5, accel=4.0)),
])
def test_lm_warp_kernel_nonstandard_worlds_equal_tile_kernel(scen, N, kw):
    """The generic (fp64) instantiation's run-time paths the product configuration never takes -- Philox motor noise,
    max_speed clamp, mass != 1, accel, agents that do not collide, heavier obstacles -- bit-identical to the tile kernels."""
    mk = lambda: BatchedFormationEnv(scen, 301, N, episode_length=6, dtype=torch.float64, seed=29, **kw)  # noqa: E731
    a, b = mk(), mk()
    a.reset(); b.reset()
    for t in range(9):
        oa, ra, da, ia = a.step_random(record_actions=True)
        with nat.options(force_tile_kernel=1):
            ob, rb, db, ib = b.step_random(record_actions=True)
        for k, (x, y) in enumerate(((oa, ob), (ra, rb), (a.pos, b.pos), (a.vel, b.vel), (a.landmarks, b.landmarks),
                                    (ia["individual_reward"], ib["individual_reward"]), (a.actions, b.actions))):
            assert torch.equal(x, y), (t, k)
        assert torch.equal(da, db)


def test_lm_warp_kernel_graph_replay_matches_eager():
    """CUDA graph of fused random-policy steps (device tick: the Philox counter lives on the device) == the same steps
    launched one by one, for the three landmark scenarios in fp32 (the STD instantiation both ways)."""
    for scen, N in (("formation_hd_partial_env", 4), ("formation_hd_partial_range_env", 5), ("formation_hd_obs_env", 4)):
        mk = lambda: BatchedFormationEnv(scen, 2050, N, episode_length=7, seed=31)  # noqa: E731
        a, b = mk(), mk()
        a.reset(); b.reset()
        g = a.capture_steps(4, fused_random=True)                    # (runs one warm-up step itself)
        g.replay(); g.replay()
        torch.cuda.synchronize()
        for _ in range(9):
            b.step_random(record_actions=True)
        for k in ("pos", "vel", "landmarks", "obs", "reward", "actions", "step_count", "ep_return"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (scen, k)
