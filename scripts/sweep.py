"""Exploration helper (not part of the product): times the fused step under CUDA-graph replay for a
grid of batch sizes / knobs.  Usage on the GPU box: python scripts/sweep.py"""
import os
import sys
import itertools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
import formation_gym  # noqa: E402


LAST = ''


def run(N, E, obs=True, steps=50, reps=4, env=None, only_step=False, warm=0.0, scenario="formation_hd_env"):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    e = formation_gym.make_batched_env(scenario, E, N, 25, write_obs=obs, seed=1)
    e.reset()
    if only_step:
        e.sample_actions()
        g = e.capture_steps(steps, policy=lambda env_: None)
    else:
        g = e.capture_steps(steps)
    g.replay(); torch.cuda.synchronize()
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    t0 = time.time()
    while time.time() - t0 < warm:                 # warm the clocks / power state
        g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    clk, pw = [], []
    while not b.query():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
        time.sleep(0.002)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (steps * reps)
    gbs = e.bytes_per_env_step() * E / us / 1e3
    global LAST
    LAST = "clk %s-%s MHz power %.0f-%.0f W (%d samples)" % (min(clk or [0]), max(clk or [0]), min(pw or [0]), max(pw or [0]), len(clk))
    for k in (env or {}):
        os.environ.pop(k, None)
    return us, gbs


if __name__ == "__main__":
    for N, E in ((243, 1024), (81, 8192), (49, 16384)):
        for mn in (48, 100000):
            us, gbs = run(N, E, True, only_step=True, reps=8, env={"FG_ROW_TMA_MIN": mn})
            print(N, E, "row_tma_min", mn, "%.1f us  %.0f GB/s" % (us, gbs), flush=True)
    os.environ["FG_FORCE_TILE_KERNEL"] = "1"
    for N, E in ((27, 65536),):
        for mn in (48, 100000):
            us, gbs = run(N, E, True, only_step=True, reps=8, env={"FG_ROW_TMA_MIN": mn})
            print("tile kernel:", N, E, "row_tma_min", mn, "%.1f us  %.0f GB/s" % (us, gbs), flush=True)
