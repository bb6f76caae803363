"""Where does the HOST time of a launch-bound step go?  cProfile of (a) the E=1 facade (configs[0]) and (b) per-step
launches of a small batch (configs[1]).  Usage: python scripts/host_profile.py"""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import numpy as np
import torch
import formation_gym


def prof(fn, n, label, top=18):
    fn(50)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(n); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("== %s: %.2f us per step" % (label, dt / n * 1e6))
    pr = cProfile.Profile(); pr.enable(); fn(n); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(top)
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:4200])


fenv = formation_gym.make_env("basic_formation_env", False, 3, 25)
fenv.seed(0); fenv.reset()


def facade(n):
    for _ in range(n):
        act_n = [sp.sample() for sp in fenv.action_space]
        _, _, done_n, _ = fenv.step(act_n)
        if np.all(done_n):
            fenv.reset()


prof(facade, 2000, "facade basic N=3 E=1 (test.py -r loop)")
fenv9 = formation_gym.make_env("formation_hd_env", False, 9, 25)
fenv9.seed(0); fenv9.reset()


def facade9(n):
    for _ in range(n):
        act_n = [sp.sample() for sp in fenv9.action_space]
        _, _, done_n, _ = fenv9.step(act_n)
        if np.all(done_n):
            fenv9.reset()


prof(facade9, 1000, "facade hd N=9 E=1", 8)
env = formation_gym.make_batched_env("formation_hd_env", 4096, 9, 25, seed=1)
env.reset()


def two(n):
    for _ in range(n):
        env.sample_actions(); env.step(env.actions)


def one(n):
    for _ in range(n):
        env.step_random(record_actions=True)


prof(two, 5000, "batched E=4096 N=9: sample_actions + step (2 launches)", 12)
prof(one, 5000, "batched E=4096 N=9: step_random(record_actions) (1 launch)", 12)
