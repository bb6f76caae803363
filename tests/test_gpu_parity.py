"""GPU parity tests: the CUDA path (through the C ABI, via BatchedFormationEnv) against
 (a) the golden fixtures frozen from the unmodified reference (tests/golden/*.npz) and
 (b) the numpy oracle on seeded random inputs.

Tolerances (BASELINE.json north_star):
  fp32 build : single step |err| <= 1e-5 on pos / vel / obs / individual reward;
               shared reward atol 1e-5 + rtol 1e-6 (|R| reaches ~900 at N=243, where one fp32 ulp
               is 6e-5 -- SURVEY.md section 7)
  fp64 build : single step <= 1e-12; 25-step trajectories <= 1e-9
"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import formation_gym  # noqa: E402
from formation_gym import _native as nat  # noqa: E402
from formation_gym.batched import BatchedFormationEnv  # noqa: E402
from oracle import mpe_oracle as mo  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32_TOL = 1e-5


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def dev(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")


def inject_hd(env, g, lm=True):
    env.pos.copy_(dev(g["pos0"], env.dtype))
    env.vel.copy_(dev(g["vel0"], env.dtype))
    env.ideal_shape.copy_(dev(g["shape"], env.dtype))
    env.ideal_vel.copy_(dev(g["ivel"], env.dtype))
    if lm and env.landmarks is not None:
        env.landmarks.copy_(dev(g["lm0"], env.dtype))
    if "step0" in g:
        env.step_count.copy_(torch.as_tensor(g["step0"], dtype=torch.int32, device="cuda"))


def maxerr(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    return float(np.max(np.abs(a - np.asarray(b, np.float64)))) if a.size else 0.0


HD_SINGLE = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "hd_n*_spread.npz"))
                   + glob.glob(os.path.join(GOLD, "hd_n*_clustered.npz")))


@pytest.mark.parametrize("track_landmarks", [True, False], ids=["tile+landmarks", "product"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("name", HD_SINGLE)
def test_hd_single_step_golden(name, dtype, track_landmarks):
    """Fixtures frozen from the unmodified reference.  track_landmarks=True also checks the landmark shift of the
    observation hook (formation_hd_env.py:40-44) and therefore runs on the tile kernel; track_landmarks=False is the
    product configuration: N = 3 / 9 / 27 take the warp-autonomous kernel k_hd_warp (the headline kernel), N = 243 the
    tile kernel with the packed pair loops."""
    g = load(name)
    E, N = g["pos0"].shape[:2]
    env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, dtype=dtype,
                              auto_reset=False, track_landmarks=track_landmarks)
    inject_hd(env, g)
    obs, rew, done, info = env.step(dev(g["act"], dtype))
    tol = F32_TOL if dtype == torch.float32 else 1e-12
    assert maxerr(env.pos, g["pos"]) <= tol
    assert maxerr(env.vel, g["vel"]) <= tol
    o = obs if "obs_rows" not in g else obs[:, torch.as_tensor(g["obs_rows"], device="cuda")]
    assert maxerr(o, g["obs"]) <= tol
    assert maxerr(info["individual_reward"], g["indiv"]) <= (tol if dtype == torch.float32 else 1e-11)
    R = g["reward"]
    rtol = 1e-6 if dtype == torch.float32 else 1e-13
    assert np.all(np.abs(rew[:, :, 0].double().cpu().numpy() - R) <= tol + rtol * np.abs(R))
    assert np.array_equal(done.cpu().numpy(), g["done"])
    assert np.array_equal(env.step_count.cpu().numpy(), g["step0"] + 1)
    if track_landmarks:
        assert maxerr(env.landmarks, g["landmarks"]) <= (1e-5 if dtype == torch.float32 else 1e-12)


@pytest.mark.parametrize("n", [3, 9, 27, 243])
def test_hd_traj25_f64(n):
    """25-step trajectories from the unmodified reference, fp64 build: <= 1e-9 (north star)."""
    g = load("hd_n%d_traj25.npz" % n)
    env = BatchedFormationEnv("formation_hd_env", 1, n, episode_length=100, dtype=torch.float64,
                              auto_reset=False)
    inject_hd(env, {k: v[None] for k, v in g.items() if k in ("pos0", "vel0", "shape", "ivel")}, lm=False)
    worst = 0.0
    for t in range(g["acts"].shape[0]):
        obs, rew, done, info = env.step(dev(g["acts"][t][None], torch.float64))
        worst = max(worst, maxerr(env.pos[0], g["pos"][t]), maxerr(env.vel[0], g["vel"][t]),
                    maxerr(info["individual_reward"][0], g["indiv"][t]))
        assert abs(float(rew[0, 0, 0]) - g["reward"][t][0]) <= 1e-9 * max(1.0, abs(g["reward"][t][0]))
    o = obs[0] if "obs_rows" not in g else obs[0][torch.as_tensor(g["obs_rows"], device="cuda")]
    worst = max(worst, maxerr(o, g["obs_last"]))
    print("N=%d 25-step fp64 max abs err %.3e" % (n, worst))
    assert worst <= 1e-9


@pytest.mark.parametrize("n", [3, 9, 27])
def test_hd_traj25_f32_reported(n):
    """fp32 over 25 steps is not a contract number (contacts amplify rounding); it must stay sane."""
    g = load("hd_n%d_traj25.npz" % n)
    env = BatchedFormationEnv("formation_hd_env", 1, n, episode_length=100, auto_reset=False)
    inject_hd(env, {k: v[None] for k, v in g.items() if k in ("pos0", "vel0", "shape", "ivel")}, lm=False)
    for t in range(g["acts"].shape[0]):
        env.step(dev(g["acts"][t][None], torch.float32))
    err = maxerr(env.pos[0], g["pos"][-1])
    print("N=%d 25-step fp32 max abs pos err %.3e" % (n, err))
    assert err <= 5e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_hd_hetero_golden(dtype):
    """per-agent mass / accel / max_speed (core.py:235-236,271-276,314-318; accel-twice quirk)."""
    g = load("hd_n9_hetero.npz")
    t = g["tweak"]
    E, N = g["pos0"].shape[:2]
    env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, dtype=dtype, auto_reset=False,
                              agent_mass=t[:, 0], agent_accel=t[:, 1], agent_max_speed=t[:, 2])
    inject_hd(env, g, lm=False)
    obs, rew, done, info = env.step(dev(g["act"], dtype))
    tol = F32_TOL if dtype == torch.float32 else 1e-12
    assert maxerr(env.pos, g["pos"]) <= tol
    assert maxerr(env.vel, g["vel"]) <= tol
    assert maxerr(obs, g["obs"]) <= tol
    assert maxerr(info["individual_reward"], g["indiv"]) <= tol * 10


def test_hd_nan_quirk():
    """Coincident agents -> NaN (core.py:312; train/README.md:194-197): same NaN pattern."""
    g = load("hd_n3_nan.npz")
    env = BatchedFormationEnv("formation_hd_env", 1, 3, episode_length=25, dtype=torch.float64,
                              auto_reset=False)
    inject_hd(env, {k: v[None] for k, v in g.items() if k in ("pos0", "vel0", "shape", "ivel")}, lm=False)
    obs, rew, done, info = env.step(dev(g["act"][None], torch.float64))
    assert np.array_equal(np.isnan(env.pos[0].cpu().numpy()), np.isnan(g["pos"]))
    assert np.array_equal(np.isnan(env.vel[0].cpu().numpy()), np.isnan(g["vel"]))
    assert np.array_equal(np.isnan(obs[0].cpu().numpy()), np.isnan(g["obs"]))
    finite = np.isfinite(g["pos"])
    assert maxerr(env.pos[0][torch.as_tensor(finite, device="cuda")], g["pos"][finite]) <= 1e-12


BASIC = ["basic_n3_spread.npz", "basic_n3_clustered.npz", "basic_n5_clustered.npz",
         "basic_n3_hetero.npz", "basic_n3_walls.npz"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("name", BASIC)
def test_basic_single_step_golden(name, dtype):
    g = load(name)
    E, N = g["pos0"].shape[:2]
    L = g["lm0"].shape[1]
    kw = {}
    if "tweak" in g:
        t = g["tweak"]
        kw = dict(agent_mass=t[:, 0], agent_accel=[None if x < 0 else x for x in t[:, 1]],
                  agent_max_speed=[None if x < 0 else x for x in t[:, 2]])
    if "walls" in g:
        kw["walls"] = [('H' if w[0] == 0 else 'V', w[1], w[2], w[3], w[4]) for w in g["walls"]]
    env = BatchedFormationEnv("basic_formation_env", E, N, episode_length=25, num_landmarks=L,
                              dtype=dtype, auto_reset=False, **kw)
    env.pos.copy_(dev(g["pos0"], dtype)); env.vel.copy_(dev(g["vel0"], dtype))
    env.landmarks.copy_(dev(g["lm0"], dtype))
    env.step_count.copy_(torch.as_tensor(g["step0"], dtype=torch.int32, device="cuda"))
    obs, rew, done, info = env.step(dev(g["act"], dtype))
    tol = F32_TOL if dtype == torch.float32 else 1e-11
    assert maxerr(env.pos, g["pos"]) <= tol
    assert maxerr(env.vel, g["vel"]) <= tol
    assert maxerr(obs, g["obs"]) <= tol
    assert maxerr(info["individual_reward"], g["indiv"]) <= tol
    assert maxerr(rew[:, :, 0], g["reward"]) <= tol * 4
    assert np.array_equal(done.cpu().numpy(), g["done"])


def test_basic_traj25_f64():
    g = load("basic_n3_traj25.npz")
    env = BatchedFormationEnv("basic_formation_env", 1, 3, episode_length=50, dtype=torch.float64,
                              auto_reset=False)
    env.pos.copy_(dev(g["pos0"][None], torch.float64)); env.vel.copy_(dev(g["vel0"][None], torch.float64))
    env.landmarks.copy_(dev(g["lm0"][None], torch.float64))
    worst = 0.0
    for t in range(g["acts"].shape[0]):
        obs, rew, done, info = env.step(dev(g["acts"][t][None], torch.float64))
        worst = max(worst, maxerr(env.pos[0], g["pos"][t]), maxerr(env.vel[0], g["vel"][t]),
                    maxerr(info["individual_reward"][0], g["indiv"][t]))
    worst = max(worst, maxerr(obs[0], g["obs_last"]))
    print("basic 25-step fp64 max abs err %.3e" % worst)
    assert worst <= 1e-9


# ------------------------------------------------------------------ oracle on seeded random batches
@pytest.mark.parametrize("E,N,spread", [(1000, 3, 0.08), (1000, 9, 0.2), (257, 9, 1.0), (64, 27, 0.35),
                                        (37, 10, 0.2), (5, 81, 0.5), (3, 243, 0.8), (2, 256, 0.9)])
def test_hd_vs_oracle_random(E, N, spread):
    """Ragged sizes (E not a multiple of the tile, N not a power of 3, N = FG_MAX_AGENTS)."""
    rng = np.random.default_rng(E * 1000 + N)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-spread, spread, (E, N, 2)))
    vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2)))
    lm = rng.uniform(-1, 1, (E, N, 2))
    shape = f32(lm - lm.mean(1, keepdims=True))
    ivel = f32(rng.uniform(-1, 1, (E, 2)))
    step0 = rng.integers(0, 25, E)
    ref = mo.hd_env_step(pos, vel, act, shape, ivel, step0, mo.WorldParams(world_length=25))
    for dtype, tol in ((torch.float32, F32_TOL), (torch.float64, 1e-11)):
        env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, dtype=dtype, auto_reset=False)
        inject_hd(env, dict(pos0=pos, vel0=vel, shape=shape, ivel=ivel, step0=step0), lm=False)
        obs, rew, done, info = env.step(dev(act, dtype))
        assert maxerr(env.pos, ref["pos"]) <= tol
        assert maxerr(env.vel, ref["vel"]) <= tol
        assert maxerr(obs, ref["obs"]) <= tol
        assert maxerr(info["individual_reward"], ref["indiv"]) <= tol * (1 if dtype == torch.float32 else 10)
        R = ref["reward"]
        assert np.all(np.abs(rew[:, 0, 0].double().cpu().numpy() - R) <= tol + 1e-6 * np.abs(R))
        assert np.array_equal(done[:, 0].cpu().numpy(), ref["done"])
        assert bool((done == done[:, :1]).all())
        assert bool((rew == rew[:, :1]).all())


@pytest.mark.parametrize("cells", [1, 0], ids=["cells", "filters"])
@pytest.mark.parametrize("E,N,scale", [(6, 32, 0.12), (5, 81, 0.12), (3, 243, 0.12), (2, 256, 0.12), (3, 243, 1.0),
                                       (4, 100, 0.3), (9, 64, 0.05)])
def test_no_obs_kernel_vs_oracle(E, N, scale, cells):
    """write_obs=False at N >= 32 runs the packed pair loops with cell lists (fg_pairs.cuh) -- the kernel the FP32
    figure is quoted on.  Against the ORACLE (not another kernel), dense clusters (scale 0.12: tens of contact partners
    and reward collisions per agent), fp32 single step: 1e-5 abs on pos / vel / individual reward, shared reward
    atol 1e-5 + rtol 1e-6; collision counts exact (they are integers inside the rewards)."""
    rng = np.random.default_rng(E * 7919 + N)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    pos = f32(rng.uniform(-scale, scale, (E, N, 2)))
    vel = f32(rng.uniform(-0.5, 0.5, (E, N, 2)))
    act = f32(rng.uniform(-1, 1, (E, N, 2)))
    lm = rng.uniform(-1, 1, (E, N, 2))
    shape = f32(lm - lm.mean(1, keepdims=True))
    ivel = f32(rng.uniform(-1, 1, (E, 2)))
    step0 = rng.integers(0, 25, E)
    ref = mo.hd_env_step(pos, vel, act, shape, ivel, step0, mo.WorldParams(world_length=25))
    with nat.options(no_cells=int(not cells)):
        env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, auto_reset=False, write_obs=False)
        inject_hd(env, dict(pos0=pos, vel0=vel, shape=shape, ivel=ivel, step0=step0), lm=False)
        obs, rew, done, info = env.step(dev(act, torch.float32))
        torch.cuda.synchronize()
    assert obs is None
    assert maxerr(env.pos, ref["pos"]) <= F32_TOL
    assert maxerr(env.vel, ref["vel"]) <= F32_TOL
    assert maxerr(info["individual_reward"], ref["indiv"]) <= F32_TOL
    R = ref["reward"]
    assert np.all(np.abs(rew[:, 0, 0].double().cpu().numpy() - R) <= F32_TOL + 1e-6 * np.abs(R))
    assert np.array_equal(done[:, 0].cpu().numpy(), ref["done"])
    # reward collisions: the per-agent count is the integer part separating indiv from the shared base term
    base = ref["indiv"].max(axis=1, keepdims=True)                       # an agent without collisions (if any)
    col_ref = np.rint(base - ref["indiv"])
    got = info["individual_reward"].double().cpu().numpy()
    col_got = np.rint(got.max(axis=1, keepdims=True) - got)
    assert np.array_equal(col_ref, col_got)
    if scale <= 0.12:
        assert col_ref.sum() > 0                                          # the dense case does exercise collisions


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_separate_entry_points_match_fused(dtype):
    """fg_world_step + fg_obs_reward == fg_step_fused: bit-exact in the fp64 build (no FMA
    contraction anywhere); the fp32 instantiations may contract differently -> 1e-6."""
    E, N = 300, 9
    a = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=3, auto_reset=False, dtype=dtype)
    b = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=3, auto_reset=False, dtype=dtype)
    a.reset(); b.reset()
    a.pos.mul_(0.15); b.pos.mul_(0.15)
    act = a.sample_actions().clone()
    a.step(act)
    b.world_step(act); b.observe()
    torch.cuda.synchronize()
    for k in ("pos", "vel", "obs", "reward", "indiv"):
        if dtype == torch.float64:
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        else:
            # a few fp32 ulps of the value (|shared reward| ~ 25 here: 1 ulp = 1.9e-6)
            ref = getattr(b, k).double().cpu().numpy()
            assert maxerr(getattr(a, k), ref) <= 1e-6 * max(1.0, float(np.abs(ref).max())), k


# ------------------------------------------------------------------ warp kernel vs tile kernel
def _run_steps(E, N, dtype, steps, force_tile, episode_length=7, u_noise=None, rollout=False, seed=11, no_std=0):
    """`steps` random-policy steps with auto-reset; FG_FORCE_TILE_KERNEL=1 routes fg_step_fused to
    the generic tile kernel (fg_kernels.cuh) instead of the warp-autonomous one (fg_warp.cuh)."""
    with nat.options(force_tile_kernel=int(force_tile), no_std_kernel=int(no_std)):
        env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=episode_length, dtype=dtype,
                                  seed=seed, auto_reset=True, u_noise=u_noise)
        env.reset()
        env.pos.mul_(0.25)                      # dense: contacts and reward-collisions do occur
        hist = []
        if rollout:
            env.rollout_random(steps)
        else:
            for _ in range(steps):
                obs, rew, done, info = env.step_random()
                hist.append((rew[:, 0, 0].clone(), done[:, 0].clone(), info["individual_reward"].clone()))
        torch.cuda.synchronize()
        out = {k: getattr(env, k).clone() for k in ("pos", "vel", "obs", "reward", "indiv", "ideal_shape",
                                                    "ideal_vel", "step_count", "ep_return", "ep_collisions",
                                                    "stats")}
        out["done"] = env.done.clone()
        return out, hist


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("E,N", [(1001, 3), (10, 3), (1000, 9), (4, 9), (257, 9), (65, 27), (1, 27),
                                 (333, 4), (7, 4), (333, 5), (100, 8), (3, 8), (65, 16), (2, 16),
                                 (200, 6), (200, 7), (40, 25), (40, 32), (1, 32)])
def test_warp_kernel_matches_tile_kernel(E, N, dtype):
    """Both kernels implement the same arithmetic in the same order: bit-exact in the fp64 build;
    fp32 may contract FMAs differently (a few ulp)."""
    a, ha = _run_steps(E, N, dtype, 16, force_tile=False)
    b, hb = _run_steps(E, N, dtype, 16, force_tile=True)
    for k in a:
        if dtype == torch.float64 or not a[k].is_floating_point():
            if k == "stats":       # atomics: summation order differs between launch geometries
                assert torch.allclose(a[k], b[k], rtol=1e-12), k
            else:
                assert torch.equal(a[k], b[k]), k
        else:
            ref = b[k].double().cpu().numpy()
            assert maxerr(a[k], ref) <= 2e-5 * max(1.0, float(np.abs(ref).max())), k
    assert float(a["stats"][0]) == 2 * E        # 16 steps, episode_length 7 -> two episode ends per env
    for (ra, da, ia), (rb, db, ib) in zip(ha, hb):
        assert torch.equal(da, db)
        if dtype == torch.float64:
            assert torch.equal(ra, rb) and torch.equal(ia, ib)


@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8, 9, 16, 25, 27, 32])
def test_warp_kernel_rollout_equals_stepwise(N):
    """n_steps random-policy steps inside ONE launch == the same steps as separate launches
    (same Philox counters), bit for bit, including the auto-resets in between."""
    # (both on the generic instantiation: single fp32 steps otherwise take the STD instantiation, whose FMA contraction
    # may differ by an ulp -- test_std_instantiation_matches_generic covers that pair)
    a, _ = _run_steps(333, N, torch.float32, 20, force_tile=False, rollout=False, no_std=1)
    b, _ = _run_steps(333, N, torch.float32, 20, force_tile=False, rollout=True, no_std=1)
    for k in a:
        if k == "stats":
            assert torch.allclose(a[k], b[k], rtol=1e-12)
        else:
            assert torch.equal(a[k], b[k]), k


def test_warp_kernel_noise_matches_tile_kernel():
    """u_noise draws come from the same Philox counters in both kernels."""
    a, _ = _run_steps(500, 9, torch.float64, 5, force_tile=False, u_noise=0.3)
    b, _ = _run_steps(500, 9, torch.float64, 5, force_tile=True, u_noise=0.3)
    for k in ("pos", "vel", "obs", "reward"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("write_obs", [True, False], ids=["obs", "noobs"])
@pytest.mark.parametrize("scen,N", [("formation_hd_env", n) for n in (3, 4, 5, 6, 7, 8, 9, 16, 25, 27, 32)]
                         + [("basic_formation_env", 3)])
def test_warp_kernel_full_wave_geometry(scen, N, write_obs):
    """Batches large enough for the full launch geometry (8- or 4-warp CTAs, one resident wave) -- the sizes bench.py
    runs -- give, env for env, the bits a small batch gives (results never depend on the launch geometry).  Guards the
    dynamic-shared-memory limit of every instantiation: N = 3 needs 31 KB per CTA, below the 48 KB default, and a limit
    left lower by the occupancy probe made exactly these launches fail."""
    EPW = 32 // N
    E_big, E_small = 2500 * EPW, 3 * EPW + 1
    outs = []
    for E in (E_big, E_small):
        env = BatchedFormationEnv(scen, E, N, episode_length=3, seed=17, auto_reset=True, write_obs=write_obs)
        env.reset()
        for _ in range(4):
            env.step_random()
        torch.cuda.synchronize()
        outs.append({k: getattr(env, k)[:E_small].clone() for k in ("pos", "vel", "reward", "indiv", "step_count")
                     + (("obs",) if write_obs else ())})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("scen,N", [("formation_hd_env", n) for n in (3, 4, 5, 6, 7, 8, 9, 16, 25, 27, 32)]
                         + [("basic_formation_env", 3)])
def test_std_instantiation_matches_generic(scen, N):
    """k_hd_warp<..., STD = true> (the standard product configuration with its run-time flag tests compiled out)
    against the generic instantiation of the same source (no_std_kernel = 1): same arithmetic in the same order;
    fp32 FMA contraction may differ between two instantiations, so a few ulp are allowed, integers are exact."""
    outs = []
    for no_std in (0, 1):
        with nat.options(no_std_kernel=no_std):
            env = BatchedFormationEnv(scen, 555, N, episode_length=5, seed=23, auto_reset=True)
            env.reset()
            env.pos.mul_(0.3)
            for _ in range(12):
                env.step_random(record_actions=True)
            torch.cuda.synchronize()
            outs.append({k: getattr(env, k).clone() for k in ("pos", "vel", "obs", "reward", "indiv", "step_count",
                                                              "ep_return", "ep_collisions", "actions", "stats")})
    a, b = outs
    for k in a:
        if not a[k].is_floating_point() or k == "actions":
            assert torch.equal(a[k], b[k]), k
        else:
            ref = b[k].double().cpu().numpy()
            assert maxerr(a[k], ref) <= 1e-5 * max(1.0, float(np.abs(ref).max())), k


def test_warp_kernel_unaligned_obs_and_no_obs():
    """The observation tensor may start on an odd 8-byte slot (head item by plain store), and
    write_obs=False must give the same state/reward."""
    E, N = 77, 9
    env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=5, auto_reset=False)
    env.reset(); env.pos.mul_(0.2)
    act = env.sample_actions().clone()
    st = {k: getattr(env, k).clone() for k in ("pos", "vel")}
    big = torch.full((E * N * env.D + 2,), -7.0, device="cuda")
    obs_u = big[2:].view(E, N, env.D)                   # 8-byte aligned, not 16
    assert obs_u.data_ptr() % 16 == 8
    b = env._make_buffers(act=act, obs=obs_u)
    env._launch_fused(b, 1, 0)
    ref_pos, ref_rew = env.pos.clone(), env.reward.clone()
    env.pos.copy_(st["pos"]); env.vel.copy_(st["vel"]); env.step_count.zero_()
    env.ep_return.zero_(); env.ep_collisions.zero_()
    obs, rew, done, info = env.step(act)
    assert torch.equal(obs, obs_u) and torch.equal(env.pos, ref_pos)
    assert float(big[0]) == -7.0 and float(big[1]) == -7.0
    env2 = BatchedFormationEnv("formation_hd_env", E, N, episode_length=25, seed=5, auto_reset=False,
                               write_obs=False)
    env2.pos.copy_(st["pos"]); env2.vel.copy_(st["vel"])
    env2.ideal_shape.copy_(env.ideal_shape); env2.ideal_vel.copy_(env.ideal_vel)
    env2.step(act)
    assert torch.equal(env2.pos, ref_pos) and torch.equal(env2.reward, ref_rew)


@pytest.mark.parametrize("N,E", [(9, 500), (27, 40), (81, 6)])
def test_graph_replay_equals_stepwise(N, E):
    """A CUDA graph of per-step launches (device-side Philox tick) replays the same trajectory as
    the same number of ordinary per-step launches -- warp kernel (N=9, 27) and tile kernel (N=81)."""
    a = BatchedFormationEnv("formation_hd_env", E, N, episode_length=6, seed=21, auto_reset=True)
    b = BatchedFormationEnv("formation_hd_env", E, N, episode_length=6, seed=21, auto_reset=True)
    a.reset(); b.reset()
    g = b.capture_steps(5)                # runs one warm-up step itself
    g.replay(); g.replay()
    for _ in range(11):
        a.step_random()
    torch.cuda.synchronize()
    for k in ("pos", "vel", "obs", "reward", "indiv", "ideal_shape", "ideal_vel", "step_count", "ep_return"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert int(b._tick_dev[0]) == 11 and int(b._tick_dev[1]) == 0
    b.use_device_tick(False)              # fold back: both envs continue identically
    a.step_random(); b.step_random()
    assert torch.equal(a.pos, b.pos)


# ------------------------------------------------------------------ packed pair loops vs scalar tile kernel
def _run_tile(E, N, steps, fast, scale, rollout=False, obs=True, seed=31):
    """fp32 tile kernel with (fast) or without (FG_NO_FAST_PAIRS=1) the packed pair loops of fg_pairs.cuh."""
    with nat.options(no_fast_pairs=int(not fast), force_fast_pairs=int(fast)):
        env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=5, seed=seed, auto_reset=True,
                                  write_obs=obs)
        env.reset()
        env.pos.mul_(scale)
        if rollout:
            env.rollout_random(steps)
        else:
            for _ in range(steps):
                env.step_random()
        torch.cuda.synchronize()
        keys = ["pos", "vel", "reward", "indiv", "ideal_shape", "ideal_vel", "step_count", "ep_return",
                "ep_collisions"] + (["obs"] if obs else [])
        out = {k: getattr(env, k).clone() for k in keys}
        out["done"] = env.done.clone()
        return out


@pytest.mark.parametrize("E,N,scale", [(7, 32, 0.3), (5, 81, 0.5), (6, 100, 1.0), (3, 243, 1.0), (2, 243, 0.12),
                                       (2, 256, 0.6), (9, 33, 0.05)])
def test_fast_pairs_match_scalar_tile_kernel(E, N, scale):
    """The group filters only select candidates; flagged pairs get the same exact tests and forces as the
    scalar loops, in the same order.  Positions/velocities agree to fp32 rounding of the force sums, the
    integer collision counts exactly.  scale 0.12 / 0.05 packs the agents densely (tens of contact partners
    per agent: the register list of near partners is flushed several times)."""
    a = _run_tile(E, N, 7, True, scale)
    b = _run_tile(E, N, 7, False, scale)
    assert torch.equal(a["ep_collisions"], b["ep_collisions"])
    assert torch.equal(a["done"], b["done"]) and torch.equal(a["step_count"], b["step_count"])
    for k in ("pos", "vel", "obs", "indiv", "reward", "ep_return", "ideal_shape", "ideal_vel"):
        ref = b[k].double().cpu().numpy()
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(a[k].double().cpu().numpy())), k
        err = np.abs(a[k].double().cpu().numpy()[fin] - ref[fin]).max() if fin.any() else 0.0
        assert err <= 2e-5 * max(1.0, float(np.abs(ref[fin]).max())), (k, err)


def test_fast_pairs_rollout_equals_stepwise():
    a = _run_tile(5, 81, 12, True, 0.4, rollout=False)
    b = _run_tile(5, 81, 12, True, 0.4, rollout=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_fast_pairs_nan_env_matches_reference_pattern():
    """Coincident agents (core.py:312) inside a large env: the NaN spreads exactly as in the scalar kernel."""
    outs = []
    for fast in (True, False):
        with nat.options(no_fast_pairs=int(not fast), force_fast_pairs=int(fast)):
            env = BatchedFormationEnv("formation_hd_env", 3, 81, episode_length=25, seed=2, auto_reset=False)
            env.reset()
            env.pos[1, 40] = env.pos[1, 7]                 # env 1: agents 7 and 40 coincide
            act = env.sample_actions().clone()
            for _ in range(3):
                env.step(act)
            torch.cuda.synchronize()
            outs.append({k: getattr(env, k).clone() for k in ("pos", "vel", "reward", "indiv", "obs")})
    a, b = outs
    assert bool(torch.isnan(a["pos"][1]).all()) and not bool(torch.isnan(a["pos"][0]).any())
    for k in a:
        assert torch.equal(torch.isnan(a[k]), torch.isnan(b[k])), k


def test_fast_pairs_far_agents_take_the_exhaustive_fallback():
    """Cell lists (fg_pairs.cuh) cover |x| < 100; an env with a farther agent is flagged and scanned exhaustively:
    same contacts, same collision counts as the scalar kernel (the far agent has a partner in contact range)."""
    outs = []
    for fast in (True, False):
        with nat.options(no_fast_pairs=int(not fast), force_fast_pairs=int(fast)):
            env = BatchedFormationEnv("formation_hd_env", 4, 243, episode_length=25, seed=5, auto_reset=False,
                                      write_obs=False)
            env.reset()
            env.pos[2] *= 0.25                                           # env 2: dense, many contacts
            env.pos[2, 11] = torch.tensor([150.0, -3.0], device="cuda")  # ... with two agents far out,
            env.pos[2, 200] = torch.tensor([150.02, -3.01], device="cuda")   # in contact AND in reward collision
            env.pos[3, 5] = torch.tensor([-1e6, 2e6], device="cuda")     # env 3: one agent very far, alone
            act = env.sample_actions().clone()
            for _ in range(2):
                env.step(act)
            torch.cuda.synchronize()
            outs.append({k: getattr(env, k).clone() for k in ("pos", "vel", "reward", "indiv", "ep_collisions")})
    a, b = outs
    assert torch.equal(a["ep_collisions"], b["ep_collisions"]) and int(a["ep_collisions"][2]) >= 2
    assert float((a["vel"][2, 11] - a["vel"][2, 200]).abs().max()) > 1.0        # the far pair pushed each other apart
    for k in ("pos", "vel", "indiv", "reward"):
        ref = b[k].double().cpu().numpy()
        err = np.abs(a[k].double().cpu().numpy() - ref).max()
        assert err <= 2e-5 * max(1.0, float(np.abs(ref).max())), (k, err)


@pytest.mark.parametrize("scen,E,N,kw", [("formation_hd_env", 1000, 5, {}), ("formation_hd_env", 77, 10, {}),
                                         ("formation_hd_env", 300, 15, {}), ("basic_formation_env", 1000, 3, {}),
                                         ("basic_formation_env", 90, 6, dict(num_landmarks=4)),
                                         ("formation_hd_partial_env", 500, 5, dict(num_obs=3)),
                                         ("formation_hd_partial_range_env", 500, 4, dict(obs_range=0.7))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_tile_image_writer_equals_flat_writer(scen, E, N, kw, dtype):
    """Observation writer OM == 3 (own row per thread staged in shared memory, one bulk store per tile) against the
    flat item loop (FG_NO_TILE_IMAGE=1): bit-identical rows, stepwise and in a rollout with auto-resets, also for an
    observation buffer that starts on an odd 8-byte slot."""
    outs = []
    for off in ("0", "1"):
        with nat.options(no_tile_image=int(off)):
            env = BatchedFormationEnv(scen, E, N, episode_length=3, seed=11, dtype=dtype, **kw)
            big = torch.zeros(E * N * env.D + 4, dtype=dtype, device="cuda")
            env.obs = big[2:2 + E * N * env.D].view(E, N, env.D)            # fp32: odd 8-byte slot of a 16-byte line
            env._bufs = env._make_buffers()
            env.reset()
            res = []
            for _ in range(4):
                env.step_random()
                res.append(env.obs.clone())
            env.rollout_random(3)
            res.append(env.obs.clone())
            assert float(big[:2].abs().sum()) == 0.0 and float(big[-2:].abs().sum()) == 0.0   # nothing outside the buffer
            outs.append(res)
    for x, y in zip(*outs):
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("E,N", [(300, 10), (70, 12), (40, 20), (33, 33), (20, 40), (7, 48), (10, 49), (9, 64), (5, 72),
                                 (4, 81), (3, 100), (2, 243), (2, 256)])
def test_chunked_row_writer_equals_row_pieces(E, N, dtype):
    """The chunked observation writer of the tile kernel -- whole rows staged 4-128 at a time, one bulk store per chunk
    (row_chunks = 2, row_min_n = 3: forced for every N) -- against the round-1 writers (row_chunks = 0, row_min_n = 48:
    per-row pieces with the static 2/3 from a shared image for N >= 48, warp-per-row plain stores / the tile image
    below): bit-identical rows, stepwise with auto-resets (early and late path) and in a rollout, also into an
    observation buffer that starts on an odd 8-byte slot."""
    outs = []
    for mode, min_n in ((0, 48), (2, 3)):
        with nat.options(row_chunks=mode, row_min_n=min_n):
            env = BatchedFormationEnv("formation_hd_env", E, N, episode_length=3, seed=19, dtype=dtype)
            big = torch.zeros(E * N * env.D + 4, dtype=dtype, device="cuda")
            env.obs = big[2:2 + E * N * env.D].view(E, N, env.D)            # fp32: odd 8-byte slot of a 16-byte line
            env._bufs = env._make_buffers()
            env.reset()
            env.pos.mul_(0.4)
            res = []
            for _ in range(5):
                env.step_random()
                res.append(env.obs.clone())
            env.rollout_random(4)
            res.append(env.obs.clone())
            assert float(big[:2].abs().sum()) == 0.0 and float(big[-2:].abs().sum()) == 0.0   # nothing outside the buffer
            outs.append(res)
    for x, y in zip(*outs):
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))


def test_misaligned_vector_buffers_are_rejected():
    """float2 / double2 access: a 4-byte-aligned obs pointer is an argument error, not a device fault."""
    from formation_gym import _native as nat
    env = BatchedFormationEnv("formation_hd_env", 8, 5, episode_length=3)
    big = torch.zeros(8 * 5 * env.D + 2, device="cuda")
    env.obs = big[1:1 + 8 * 5 * env.D].view(8, 5, env.D)
    env._bufs = env._make_buffers()
    with pytest.raises(nat.NativeError, match="aligned"):
        env.step_random()


def _run_basic(E, dtype, steps, force_tile, rollout=False, u_noise=None, no_std=0):
    with nat.options(force_tile_kernel=int(force_tile), no_std_kernel=int(no_std)):
        env = BatchedFormationEnv("basic_formation_env", E, 3, episode_length=7, dtype=dtype, seed=13,
                                  auto_reset=True, u_noise=u_noise)
        env.reset()
        env.pos.mul_(0.3)                       # agents of size 0.1: contacts and collisions occur
        hist = []
        if rollout:
            env.rollout_random(steps)
        else:
            for _ in range(steps):
                obs, rew, done, info = env.step_random()
                hist.append((rew[:, 0, 0].clone(), done[:, 0].clone(), info["individual_reward"].clone()))
        torch.cuda.synchronize()
        out = {k: getattr(env, k).clone() for k in ("pos", "vel", "obs", "reward", "indiv", "landmarks", "step_count",
                                                    "ep_return", "ep_collisions", "stats")}
        out["done"] = env.done.clone()
        return out, hist


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("E", [1, 10, 1001])
def test_basic_warp_kernel_matches_tile_kernel(E, dtype):
    """basic_formation_env (3 agents / 3 landmarks) on the warp-autonomous kernel (fg_warp.cuh, SCN = basic) against
    the tile kernel: bit-exact in fp64 over 16 steps with two auto-resets (landmarks redrawn), a few ulp in fp32;
    the in-kernel rollout equals the stepwise one; motor noise uses the same Philox counters."""
    a, ha = _run_basic(E, dtype, 16, force_tile=False)
    b, hb = _run_basic(E, dtype, 16, force_tile=True)
    for k in a:
        if dtype == torch.float64 or not a[k].is_floating_point():
            if k == "stats":
                assert torch.allclose(a[k], b[k], rtol=1e-12), k
            else:
                assert torch.equal(a[k], b[k]), k
        else:
            ref = b[k].double().cpu().numpy()
            assert maxerr(a[k], ref) <= 2e-5 * max(1.0, float(np.abs(ref).max())), k
    assert float(a["stats"][0]) == 2 * E
    assert int(a["ep_collisions"].sum()) > 0 or E < 10
    for (ra, da, ia), (rb, db, ib) in zip(ha, hb):
        assert torch.equal(da, db)
        if dtype == torch.float64:
            assert torch.equal(ra, rb) and torch.equal(ia, ib)
    # in-kernel rollout == stepwise, bit for bit, on ONE instantiation (single fp32 steps otherwise take the STD one)
    a0, _ = _run_basic(E, dtype, 16, force_tile=False, no_std=1)
    c, _ = _run_basic(E, dtype, 16, force_tile=False, rollout=True, no_std=1)
    for k in a0:
        assert torch.allclose(a0[k], c[k], rtol=1e-12) if k == "stats" else torch.equal(a0[k], c[k]), k
    if dtype == torch.float64:
        n1, _ = _run_basic(E, dtype, 5, force_tile=False, u_noise=0.2)
        n2, _ = _run_basic(E, dtype, 5, force_tile=True, u_noise=0.2)
        for k in ("pos", "vel", "obs", "reward"):
            assert torch.equal(n1[k], n2[k]), k
