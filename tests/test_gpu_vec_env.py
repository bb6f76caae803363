"""VecEnv adapter: interface and auto-reset semantics of the reference's
SubprocVecEnv / DummyVecEnv (train/maddpg-v2/utils/env_wrappers.py:9-128) over the CUDA env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import formation_gym  # noqa: E402
from formation_gym.batched import BatchedFormationEnv  # noqa: E402
from oracle import mpe_oracle as mo  # noqa: E402


def test_vec_env_interface_numpy():
    E, N, T = 6, 9, 4
    venv = formation_gym.make_vec_env("formation_hd_env", E, N, episode_length=T, seed=3)
    assert venv.num_envs == E and venv.num_agents == N
    assert len(venv.action_space) == N and venv.action_space[0].shape == (2,)
    assert venv.observation_space[0].shape == (6 * N,)
    assert venv.share_observation_space[0].shape == (6 * N * N,)
    assert venv.agent_types == ['agent'] * N
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (E, N, 6 * N) and obs.dtype == np.float32
    assert venv.share_obs(obs).shape == (E, N * 6 * N)
    for t in range(1, 2 * T + 1):
        actions = np.stack([[sp.sample() for sp in venv.action_space] for _ in range(E)])
        venv.step_async(actions)
        with pytest.raises(RuntimeError):
            venv.step_async(actions)
        obs, rews, dones, infos = venv.step_wait()
        assert obs.shape == (E, N, 6 * N) and rews.shape == (E, N, 1) and dones.shape == (E, N)
        assert dones.dtype == np.bool_
        assert bool(dones.all()) == (t % T == 0) and bool(dones.any()) == (t % T == 0)
        assert len(infos) == E and len(infos[0]) == N
        r = infos[2][5]['individual_reward']
        assert isinstance(r, float)
        # shared reward = sum of the individual rewards (environment.py:136-138)
        assert abs(rews[2, 0, 0] - sum(d['individual_reward'] for d in infos[2])) <= 1e-4 * max(1, abs(rews[2, 0, 0]))
    with pytest.raises(RuntimeError):
        venv.step_wait()
    venv.close()
    assert venv.closed


def test_vec_env_matches_batched_env_and_oracle():
    """Same seed -> same trajectory as BatchedFormationEnv; one step checked against the numpy oracle."""
    E, N = 33, 9
    venv = formation_gym.make_vec_env("formation_hd_env", E, N, episode_length=25, seed=7)
    ref = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=7)
    o1 = venv.reset().copy()
    o2 = ref.reset()
    assert np.array_equal(o1, o2.cpu().numpy())
    st = {k: getattr(ref, k).double().cpu().numpy() for k in ("pos", "vel", "ideal_shape", "ideal_vel")}
    rng = np.random.default_rng(0)
    act = rng.uniform(-1, 1, (E, N, 2)).astype(np.float32)
    obs, rews, dones, infos = venv.step(act)
    obs2, rew2, done2, info2 = ref.step(torch.as_tensor(act, device="cuda"))
    assert np.array_equal(obs, obs2.cpu().numpy()) and np.array_equal(rews, rew2.cpu().numpy())
    assert np.array_equal(infos.individual_reward, info2["individual_reward"].cpu().numpy())
    want = mo.hd_env_step(st["pos"], st["vel"], act.astype(np.float64), st["ideal_shape"], st["ideal_vel"],
                          np.zeros(E, np.int64), mo.WorldParams(world_length=25))
    assert np.abs(obs - want["obs"]).max() <= 1e-5
    assert np.abs(infos.individual_reward - want["indiv"]).max() <= 1e-5


@pytest.mark.parametrize("scen,N,kw", [("formation_hd_env", 9, {}), ("formation_hd_env", 27, {}), ("formation_hd_env", 40, {}),
                                       ("formation_hd_env", 9, dict(dtype=torch.float64)),
                                       ("formation_hd_env", 5, dict(auto_reset=False)),
                                       ("basic_formation_env", 3, {}), ("formation_hd_partial_env", 5, {}),
                                       ("formation_hd_obs_env", 4, {}), ("formation_hd_env", 6, dict(silent=False))])
def test_vec_env_host_arrays_stay_byte_identical(scen, N, kw):
    """to_numpy=True ships only the dynamic prefix of every observation row on ordinary steps (fg_obs_to_host mode 1)
    and whole rows on the steps that end the episodes; the persistent pinned host array must equal the device tensor
    byte for byte after EVERY step, across auto-resets (new ideal shape / ideal velocity / landmarks)."""
    E, T = 37, 4
    venv = formation_gym.make_vec_env(scen, E, N, episode_length=T, seed=5, **kw)
    env = venv.env
    obs = venv.reset()
    assert np.array_equal(obs, env.obs.cpu().numpy())
    rng = np.random.default_rng(1)
    for t in range(1, 3 * T + 2):
        act = rng.uniform(-1, 1, (E, N, env.act_dim)).astype(obs.dtype)
        obs, rews, dones, infos = venv.step(act)
        assert obs.tobytes() == env.obs.cpu().numpy().tobytes(), t
        assert not obs.flags.writeable                     # incremental refresh: in-place edits are refused, not lost
        assert np.array_equal(rews, env.reward.cpu().numpy()) and np.array_equal(dones, env.done.cpu().numpy())
        if kw.get("auto_reset", True):
            assert bool(dones.all()) == (t % T == 0)
    venv.close()


def test_vec_env_auto_reset_returns_reset_obs_with_terminal_reward():
    """worker(): `if all(done): ob = env.reset()` -- terminal reward/done, RESET observation
    (env_wrappers.py:14-18)."""
    E, N, T = 5, 3, 3
    venv = formation_gym.make_vec_env("formation_hd_env", E, N, episode_length=T, seed=11, to_numpy=False)
    venv.reset()
    for _ in range(T):
        obs, rews, dones, infos = venv.step(venv.sample_actions())
    assert bool(dones.all())
    env = venv.env
    assert int(env.step_count.abs().sum()) == 0                      # current_step = 0 after the reset
    assert float(env.vel.abs().max()) == 0.0                         # reset_world: p_vel = 0
    obs_kept = obs.clone()
    assert torch.equal(env.observe(), obs_kept)                      # the returned obs IS the reset state's obs
    assert float(rews.abs().min()) > 0.0                             # terminal (non-reset) reward came through
    # device mode returns CUDA tensors without a host round trip
    assert obs.is_cuda and rews.is_cuda and dones.is_cuda and infos.individual_reward.is_cuda


def test_vec_env_pinned_action_buffer_and_basic_scenario():
    venv = formation_gym.make_vec_env("basic_formation_env", 4, 3, episode_length=5, seed=1)
    obs = venv.reset()
    assert obs.shape == (4, 3, 18)
    buf = venv.action_buffer
    buf[...] = 0.25
    obs, rews, dones, infos = venv.step(buf)
    assert obs.shape == (4, 3, 18) and rews.shape == (4, 3, 1)
    assert np.isfinite(obs).all() and np.isfinite(rews).all()


def test_render_bridge_batched():
    """BatchedFormationEnv.render / CudaVecEnv.render: one env's state copied to the host and rasterised."""
    import formation_gym
    env = formation_gym.make_batched_env("formation_hd_obs_env", 8, 4, 50, seed=2)
    env.reset()
    env.pos[3] = torch.tensor([[-0.5, 0.0], [0.5, 0.0], [0.0, 0.5], [0.0, -0.5]], device="cuda")
    frame = env.render(3)
    assert frame.shape == (700, 700, 3) and frame.dtype == np.uint8
    px = frame[350, int(350 + 0.5 * 175)].astype(int)                 # agent 1: half-transparent blue on white
    assert np.abs(px - (0.5 * 255 + 0.5 * 255 * np.array([0.35, 0.35, 0.85]))).max() <= 2
    venv = formation_gym.make_vec_env("formation_hd_env", 4, 9, 25)
    venv.reset()
    assert venv.render('rgb_array').shape == (700, 700, 3)
    venv.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("N,E", [(9, 301), (40, 17), (3, 1000)])
def test_obs_to_host_modes(N, E, dtype):
    """fg_obs_to_host: every mode leaves exactly the bytes it promises in the pinned host array -- mode 0 the whole
    tensor, mode 1 the dynamic prefix of every row, mode 2 the prefix plus WHOLE rows where done != 0 (device-side
    flags), mode 3 the packed prefixes -- and nothing else (the rest of the host array keeps its old contents)."""
    import ctypes as C
    from formation_gym import _native as nat
    lib = nat.load()
    row_items, dyn = 3 * N, N
    obs = torch.randn(E, N, 6 * N, device="cuda", dtype=dtype)
    done = torch.zeros(E, N, dtype=torch.uint8, device="cuda")
    done[::5] = 1
    stage_d = torch.empty(E * N, 2 * dyn, device="cuda", dtype=dtype)
    isz = 8 if dtype == torch.float32 else 16
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    oc = obs.cpu()

    def run(mode, dst, dn=None):
        nat.check(lib.fg_obs_to_host(obs.data_ptr(), dst.data_ptr(), None if dn is None else dn.data_ptr(),
                                     stage_d.data_ptr(), E, N, row_items, dyn, isz, mode, st), "fg_obs_to_host")
        torch.cuda.synchronize()
    host = torch.full((E, N, 6 * N), -3.0, dtype=dtype).pin_memory()
    run(0, host)
    assert torch.equal(host, oc)
    for mode in (1, 2):
        host.fill_(-3.0)
        run(mode, host)
        assert torch.equal(host[:, :, :2 * N], oc[:, :, :2 * N]) and bool((host[:, :, 2 * N:] == -3.0).all())
    host.fill_(-3.0)
    run(2, host, done)
    assert torch.equal(host[::5], oc[::5])
    keep = torch.ones(E, dtype=torch.bool); keep[::5] = False
    assert torch.equal(host[keep][:, :, :2 * N], oc[keep][:, :, :2 * N]) and bool((host[keep][:, :, 2 * N:] == -3.0).all())
    packed = torch.empty(E * N, 2 * dyn, dtype=dtype).pin_memory()
    run(3, packed)
    assert torch.equal(packed, oc.view(E * N, 6 * N)[:, :2 * N])
    # mode 4 (whole 64-byte host lines): relies on the static part of the host array already holding the device values
    if (E * N * row_items * isz) % 16 == 0:
        host.copy_(oc); host[:, :, :2 * N] = -3.0
        run(4, host)
        assert torch.equal(host, oc)
        host.copy_(oc); host[:, :, :2 * N] = -3.0; host[::5] = -3.0
        run(4, host, done)
        assert torch.equal(host, oc)
    else:
        assert lib.fg_obs_to_host(obs.data_ptr(), host.data_ptr(), None, None, E, N, row_items, dyn, isz, 4, st) == -1
    assert lib.fg_obs_to_host(obs.data_ptr(), host.data_ptr(), None, None, E, N, row_items, row_items + 1, isz, 1, st) == -1
    assert lib.fg_obs_to_host(obs.data_ptr(), host.data_ptr(), None, None, E, N, row_items, dyn, 12, 1, st) == -1
