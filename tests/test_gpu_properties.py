"""Size-independent properties of the step path, checked at small sizes and at the FULL sizes
of BASELINE.json's configs, where the numpy oracle would take minutes: momentum symmetry of the contact force
(core.py:314-318), far pairs contributing exactly nothing, permutation equivariance over agents, independence of the
results from how envs are sharded over GPUs (Philox keyed by the global env id), noise and reset statistics
(core.py:232-233, formation_hd_env.py:77-95), and the observation layout (formation_hd_env.py:52-59) recomputed with
plain torch ops from the state tensors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from formation_gym.batched import BatchedFormationEnv  # noqa: E402


def _hd(E, N, dtype=torch.float64, **kw):
    kw.setdefault("auto_reset", False)
    env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, dtype=dtype, seed=4, **kw)
    env.reset()
    return env


def test_contact_forces_conserve_momentum():
    """f_a = +f, f_b = -f for equal masses (core.py:314-318): with zero actions and zero velocities the velocities
    after one step sum to zero in every env, however many agents are in contact."""
    env = _hd(64, 27)
    env.pos.mul_(0.1)                                    # dense: many contacts
    env.vel.zero_()
    env.step(torch.zeros_like(env.actions))
    assert float(env.vel.abs().max()) > 1.0               # contacts did push
    tot = env.vel.sum(1)
    assert float(tot.abs().max()) <= 1e-11 * float(env.vel.abs().max()) * 27


def test_far_pairs_contribute_exactly_nothing():
    """Agents farther apart than the contact range: v' = v (1 - damping) + 5 a dt, p' = p + v' dt, bit for bit
    (core.py:268-277; the softplus underflows to an exact 0 in the reference, the kernel skips the pair)."""
    env = _hd(32, 9)
    grid = torch.stack(torch.meshgrid(torch.arange(3.), torch.arange(3.), indexing="ij"), -1).reshape(9, 2)
    env.pos.copy_((grid * 0.9 - 0.9).to(env.pos).expand(32, 9, 2) + 0.01 * env.pos)
    v0 = torch.rand_like(env.vel) - 0.5
    env.vel.copy_(v0)
    p0 = env.pos.clone()
    act = env.sample_actions().clone()
    env.step(act)
    v1 = v0 * 0.75 + (act * 5.0) * 0.1
    assert torch.equal(env.vel, v1) and torch.equal(env.pos, p0 + v1 * 0.1)


@pytest.mark.parametrize("N", [9, 40])
def test_permutation_equivariance(N):
    """Relabelling the agents (and the rows of the ideal shape with them) permutes positions, velocities and
    individual rewards; the formation and velocity terms and the collision total do not change."""
    a = _hd(16, N)
    a.pos.mul_(0.3)
    b = _hd(16, N)
    perm = torch.randperm(N, device="cuda")
    for k in ("pos", "vel", "ideal_shape"):
        getattr(b, k).copy_(getattr(a, k)[:, perm])
    b.ideal_vel.copy_(a.ideal_vel)
    act = a.sample_actions().clone()
    _, ra, _, ia = a.step(act)
    _, rb, _, ib = b.step(act[:, perm].contiguous())
    assert float((b.pos - a.pos[:, perm]).abs().max()) <= 1e-12          # force sums in a different order
    assert float((b.vel - a.vel[:, perm]).abs().max()) <= 1e-11
    assert float((ib["individual_reward"] - ia["individual_reward"][:, perm]).abs().max()) <= 1e-12
    assert float((rb - ra).abs().max()) <= 1e-10


@pytest.mark.parametrize("scen,N", [("formation_hd_env", 9), ("formation_hd_env", 27), ("formation_hd_env", 81),
                                    ("formation_hd_env", 16), ("formation_hd_env", 100),
                                    ("basic_formation_env", 3), ("formation_hd_obs_env", 4)])
def test_sharding_invariance(scen, N):
    """One batch of E envs == two shards with env_offset (the multi-GPU layout): identical states, observations and
    statistics after a random-policy rollout with auto-resets and motor noise."""
    E, T = 1000, 12
    kw = dict(episode_length=5, seed=9, u_noise=0.05, dtype=torch.float32)
    full = BatchedFormationEnv(scen, E, N, **kw)
    lo = BatchedFormationEnv(scen, 400, N, env_offset=0, **kw)
    hi = BatchedFormationEnv(scen, 600, N, env_offset=400, **kw)
    for env in (full, lo, hi):
        env.reset()
        for _ in range(T):
            env.step_random()
    for k in ("pos", "vel", "obs", "reward", "indiv", "step_count", "ep_return"):
        both = torch.cat([getattr(lo, k), getattr(hi, k)], 0)
        assert torch.equal(getattr(full, k), both), k
    assert torch.allclose(full.stats, lo.stats + hi.stats, rtol=1e-12, atol=0)


def test_motor_noise_statistics():
    """core.py:232-233: F += randn(2) * u_noise.  Far-apart agents at rest with zero actions: v' = noise * dt."""
    env = _hd(4096, 9, dtype=torch.float32, u_noise=0.1)
    grid = torch.stack(torch.meshgrid(torch.arange(3.), torch.arange(3.), indexing="ij"), -1).reshape(9, 2)
    env.pos.copy_(grid.to(env.pos).expand(4096, 9, 2))
    env.vel.zero_()
    env.step(torch.zeros_like(env.actions))
    n = env.vel.double().flatten() / 0.1
    assert abs(float(n.mean())) < 4 * 0.1 / np.sqrt(n.numel())
    assert abs(float(n.std()) - 0.1) < 0.002
    assert abs(float((n ** 4).mean() / n.var() ** 2) - 3.0) < 0.1         # Gaussian kurtosis
    # the two components and different agents are uncorrelated
    v = env.vel.double() / 0.1
    assert abs(float((v[..., 0] * v[..., 1]).mean())) < 5e-4
    assert abs(float((v[:, 0, 0] * v[:, 1, 0]).mean())) < 2e-3


def test_reset_statistics():
    """formation_hd_env.py:77-95: agents and landmarks ~ U(-1,1)^2, ideal_shape = landmarks - mean, ideal_vel ~
    U(-1,1)^2, velocities 0; different envs / agents / purposes draw independent numbers."""
    env = _hd(65536, 9, dtype=torch.float32)
    p = env.pos.double()
    assert float(p.abs().max()) <= 1.0 and float(env.vel.abs().max()) == 0.0
    assert abs(float(p.mean())) < 3e-3 and abs(float(p.var()) - 1 / 3) < 3e-3
    assert float(env.ideal_shape.double().mean(1).abs().max()) < 1e-6    # centred
    # variance of a centred uniform sample: (1 - 1/N) / 3
    assert abs(float(env.ideal_shape.double().var(unbiased=False)) - (8 / 9) / 3) < 3e-3
    iv = env.ideal_vel.double()
    assert abs(float(iv.mean())) < 6e-3 and abs(float(iv.var()) - 1 / 3) < 6e-3
    assert abs(float((p[:, 0, 0] * p[:, 1, 0]).mean())) < 6e-3 and abs(float((p[:, 0, 0] * iv[:, 0]).mean())) < 6e-3
    assert abs(float((p[:-1, 0, 0] * p[1:, 0, 0]).mean())) < 6e-3


def _check_hd_outputs(env, act, p0, v0):
    """Everything an env.step returns, recomputed with plain torch ops from the state tensors (exact where the
    arithmetic is a single rounded operation)."""
    E, N = env.E, env.N
    obs, rew, done, info = env.obs, env.reward, env.done, {"individual_reward": env.indiv}
    pos, vel = env.pos, env.vel
    assert torch.equal(obs[:, :, 0:2], vel)                                          # p_vel
    idx = torch.arange(N, device="cuda")
    others = torch.stack([torch.cat([idx[:i], idx[i + 1:]]) for i in range(N)])      # [N, N-1]
    for i in range(0, N, max(1, N // 9)):                                            # sampled rows (memory)
        rel = pos[:, others[i]] - pos[:, i:i + 1]
        assert torch.equal(obs[:, i, 2:2 * N].reshape(E, N - 1, 2), rel)             # other_pos
        assert float(obs[:, i, 2 * N:4 * N - 2].abs().max()) == 0.0                  # comm of silent agents
        assert torch.equal(obs[:, i, 4 * N - 2:6 * N - 2].reshape(E, N, 2), env.ideal_shape)
        assert torch.equal(obs[:, i, 6 * N - 2:], env.ideal_vel)
    # physics of agents that had nobody within the contact range: exact
    d = torch.cdist(p0, p0) + 10.0 * torch.eye(N, device="cuda")
    free = d.min(2).values > 0.09
    v1 = v0 * 0.75 + (act * 5.0) * 0.1
    # (fp32 build: the kernel may contract v * keep + F * dt into an FMA, hence one ulp)
    assert float((vel[free] - v1[free]).abs().max()) <= 2e-7 and float((pos[free] - (p0 + v1 * 0.1)[free]).abs().max()) <= 5e-7
    assert float(free.float().mean()) > 0.1
    # reward: shared sum broadcast, individual rewards = base - collisions
    assert torch.equal(rew[:, :, 0], rew[:, :1, 0].expand(E, N))
    dn = torch.cdist(pos.double(), pos.double()) + 10.0 * torch.eye(N, device="cuda", dtype=torch.float64)
    margin = (dn - 0.03).abs().min(2).values > 1e-6                                  # away from the threshold
    col = (dn < 0.03).sum(2)
    indiv = info["individual_reward"].double()
    base = (indiv + col)                                                             # same for all agents of an env
    ok = margin.all(1)
    assert float((base[ok] - base[ok][:, :1]).abs().max()) <= 2e-5
    C = pos.double() - pos.double().mean(1, keepdim=True)
    S = env.ideal_shape.double()
    d2 = torch.cdist(C, S)
    form = torch.maximum(d2.min(2).values.max(1).values, d2.min(1).values.max(1).values)
    velr = (env.ideal_vel.double() - vel.double().mean(1)).norm(dim=1)
    assert float((base[ok][:, 0] - (-form - velr)[ok]).abs().max()) <= 2e-5
    tot = (indiv.sum(1) - rew[:, 0, 0].double()).abs()
    assert bool((tot <= 1e-5 + 1e-6 * rew[:, 0, 0].double().abs()).all())
    assert torch.equal(done, (env.step_count >= 25)[:, None].expand(E, N)) or bool(env.auto_reset)


@pytest.mark.parametrize("N,E", [(9, 131072), (27, 65536), (3, 1048576), (243, 1024), (9, 4096)])
def test_full_size_step_properties(N, E):
    """BASELINE.json's configs at FULL size (fp32): one step on given actions, outputs recomputed from the state
    tensors with torch."""
    env = _hd(E, N, dtype=torch.float32)
    if N == 243:
        env.pos.mul_(1.0)
    p0, v0 = env.pos.clone(), (torch.rand_like(env.vel) - 0.5)
    env.vel.copy_(v0)
    act = env.sample_actions().clone()
    env.step(act)
    _check_hd_outputs(env, act, p0, v0)


@pytest.mark.parametrize("N", [40, 81, 243])
def test_sharding_invariance_without_observations(N):
    """State + reward only (write_obs=False: the packed pair loops with cell lists from N = 32 up): sharding the batch
    does not change a bit either."""
    E, T = 300, 8
    kw = dict(episode_length=5, seed=3, dtype=torch.float32, write_obs=False)
    full = BatchedFormationEnv("formation_hd_env", E, N, **kw)
    lo = BatchedFormationEnv("formation_hd_env", 100, N, env_offset=0, **kw)
    hi = BatchedFormationEnv("formation_hd_env", 200, N, env_offset=100, **kw)
    for env in (full, lo, hi):
        env.reset()
        env.pos.mul_(0.5)                       # denser: contacts and reward collisions
        for _ in range(T):
            env.step_random()
    for k in ("pos", "vel", "reward", "indiv", "step_count", "ep_return", "ep_collisions"):
        assert torch.equal(getattr(full, k), torch.cat([getattr(lo, k), getattr(hi, k)], 0)), k


@pytest.mark.parametrize("scen,N", [("formation_hd_env", 9), ("formation_hd_env", 40), ("basic_formation_env", 3),
                                    ("formation_hd_partial_env", 5), ("formation_hd_obs_env", 4)])
def test_masked_reset_touches_only_masked_envs(scen, N):
    """fg_reset with a mask (env.reset of SOME envs): masked envs get a fresh reset_world state and step 0, the
    others keep every bit; two different ticks draw different states."""
    E = 257
    env = BatchedFormationEnv(scen, E, N, episode_length=25, seed=6, auto_reset=False)
    env.reset()
    for _ in range(3):
        env.step_random()
    keys = [k for k in ("pos", "vel", "landmarks", "landmark_vel", "ideal_shape", "ideal_vel", "step_count")
            if getattr(env, k, None) is not None]
    before = {k: getattr(env, k).clone() for k in keys}
    mask = (torch.arange(E, device="cuda") % 3 == 1)
    env.reset(mask)
    for k in keys:
        a, b = getattr(env, k), before[k]
        assert torch.equal(a[~mask], b[~mask]), k
        if k in ("pos", "step_count"):
            assert not torch.equal(a[mask], b[mask]), k
    assert int(env.step_count[mask].abs().sum()) == 0 and int(env.step_count[~mask].min()) == 3
    assert float(env.vel[mask].abs().max()) == 0.0 and float(env.pos[mask].abs().max()) <= 1.0
    first = env.pos[mask].clone()
    env.reset(mask)
    assert not torch.equal(env.pos[mask], first)


@pytest.mark.parametrize("N", [3, 9, 40])
def test_non_silent_agents_communicate(N):
    """agent.silent = False (core.py:279-286): actions carry [u, c], the comm state becomes action.c and every agent
    observes the others' utterances in agent order (formation_hd_env.py:50-57).  Goes through the tile kernel."""
    E = 50
    env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=2, auto_reset=False, silent=False,
                              dtype=torch.float64)
    env.reset()
    assert env.act_dim == 4
    act = torch.rand(E, N, 4, device="cuda", dtype=torch.float64) * 2 - 1
    sil = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=2, auto_reset=False, dtype=torch.float64)
    sil.reset()
    obs, rew, done, info = env.step(act)
    obs_s, rew_s, _, _ = sil.step(act[..., :2].contiguous())
    assert torch.equal(env.comm, act[..., 2:])
    assert torch.equal(env.pos, sil.pos) and torch.equal(rew, rew_s)          # comm does not touch the physics
    for i in range(N):
        others = [j for j in range(N) if j != i]
        assert torch.equal(obs[:, i, 2 * N:4 * N - 2].reshape(E, N - 1, 2), act[:, others, 2:])
        assert torch.equal(obs[:, i, :2 * N], obs_s[:, i, :2 * N]) and torch.equal(obs[:, i, 4 * N - 2:], obs_s[:, i, 4 * N - 2:])
    # comm noise (core.py:284-285): c = action.c + N(0, c_noise)
    noisy = BatchedFormationEnv("formation_hd_env", 4096, 3, episode_length=25, seed=2, auto_reset=False, silent=False,
                                c_noise=0.05)
    noisy.reset()
    a4 = torch.zeros(4096, 3, 4, device="cuda")
    noisy.step(a4)
    c = noisy.comm.double().flatten()
    assert abs(float(c.mean())) < 2e-3 and abs(float(c.std()) - 0.05) < 2e-3


@pytest.mark.parametrize("scen,N,E", [("formation_hd_env", 9, 777), ("formation_hd_env", 27, 50), ("formation_hd_env", 81, 9),
                                      ("basic_formation_env", 3, 1000), ("basic_formation_env", 5, 300)])
def test_fused_random_policy_records_the_same_actions(scen, N, E):
    """step_random(record_actions=True) (fg_step_fused random_actions = 2: the policy drawn AND recorded inside the step
    kernel, one launch) == sample_actions() + step(actions) (two launches): same actions, bit-identical results,
    across auto-resets; warp kernel (N = 9, 27, basic 3) and tile kernel (N = 81, basic 5)."""
    a = BatchedFormationEnv(scen, E, N, episode_length=4, seed=9, auto_reset=True)
    b = BatchedFormationEnv(scen, E, N, episode_length=4, seed=9, auto_reset=True)
    a.reset(); b.reset()
    for _ in range(9):
        a.step_random(record_actions=True)
        b.sample_actions(); b.step(b.actions)
        assert torch.equal(a.actions, b.actions)
    for k in ("pos", "vel", "obs", "reward", "indiv", "step_count", "ep_return"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert float(a.actions.abs().max()) <= 1.0 and float(a.actions.abs().max()) > 0.9


@pytest.mark.parametrize("scen,N", [("formation_hd_env", 9), ("formation_hd_env", 40), ("formation_hd_env", 243),
                                    ("basic_formation_env", 3), ("basic_formation_env", 6), ("formation_hd_partial_env", 5),
                                    ("formation_hd_obs_env", 4)])
def test_nan_flag_marks_exactly_the_failed_envs(scen, N):
    """fg_buffers.nan_flag: sticky per-env flag of the reference's documented failure mode (coincident agents ->
    delta_pos / dist = 0/0, core.py:312; train/README.md:194-197).  Only the env with two coincident agents is flagged;
    the flag survives further steps, an auto-reset heals the env but not the flag, reset() clears it."""
    E = 12
    env = BatchedFormationEnv(scen, E, N, episode_length=3, seed=3, auto_reset=True)
    env.reset()
    assert int(env.nan_flag.sum()) == 0
    env.step_random()
    assert int(env.nan_flag.sum()) == 0 and env.nan_envs().numel() == 0
    env.pos[5, 1] = env.pos[5, 0]                                   # env 5: agents 0 and 1 coincide
    env.step_random()
    torch.cuda.synchronize()
    assert env.nan_envs().tolist() == [5]
    assert bool(torch.isnan(env.reward[5]).all()) and not bool(torch.isnan(env.reward[:5]).any())
    for _ in range(3):                                              # episode ends -> env 5 is reset and healthy again
        env.step_random()
    assert not bool(torch.isnan(env.pos).any())
    assert env.nan_envs(clear=True).tolist() == [5]                 # sticky until cleared
    assert env.nan_envs().numel() == 0


@pytest.mark.parametrize("scen,N,E", [("formation_hd_env", 9, 3000), ("formation_hd_env", 27, 200), ("basic_formation_env", 3, 5000)])
def test_l2_prefetch_does_not_change_results(scen, N, E):
    """The warp kernel's prefetch.global.L2 of the state two spans ahead (on by default only for batches whose state
    exceeds ~1/3 of L2) is a pure hint: forcing it on / off gives bit-identical trajectories."""
    from formation_gym import _native as nat
    outs = []
    for mode in (0, 2):
        with nat.options(l2_prefetch=mode):
            env = BatchedFormationEnv(scen, E, N, episode_length=4, seed=8, auto_reset=True)
            env.reset()
            for _ in range(6):
                env.step_random()
            torch.cuda.synchronize()
            outs.append({k: getattr(env, k).clone() for k in ("pos", "vel", "obs", "reward", "step_count")})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
