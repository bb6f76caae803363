"""Quick timing of the fused step kernel alone (CUDA graph of 5 steps) for a list of configs.
Usage: [FG_SETTLE=seconds] python scripts/quick_cfg.py "scen N E [obs] [opt=val,...|-] [envkw=val,...]" ...
FG_SETTLE > 0 replays the graph for that long first, so that the timing runs at the clock the box holds under its
power cap (as bench.py's timed region does) instead of the burst clock."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
import formation_gym  # noqa: E402
from formation_gym import _native as nat  # noqa: E402

PEAK = 6547.8
for cfg in sys.argv[1:]:
    f = cfg.split()
    scen, N, E = f[0], int(f[1]), int(f[2])
    obs = (f[3] != "0") if len(f) > 3 else True
    opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in f[4].split(",")) if len(f) > 4 and f[4] != "-" else {}
    envkw = dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in f[5].split(",")) if len(f) > 5 else {}
    with nat.options(**opts):
        env = formation_gym.make_batched_env(scen, E, N, 25, write_obs=obs, seed=1, **envkw)
        env.reset()
        env.sample_actions()
        g = env.capture_steps(5, policy=lambda e_: None)
        g.replay(); torch.cuda.synchronize()
        settle = float(os.environ.get("FG_SETTLE", "0"))
        if settle > 0:
            import time
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < settle:
                for _ in range(20):
                    g.replay()
                torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 50)
        gbs = env.bytes_per_env_step() * E / (best * 1e-3) / 1e9
        print("%-28s N=%3d E=%8d obs=%d %-24s %9.2f us  %7.0f GB/s  frac %.3f" % (scen, N, E, obs, opts, best * 1e3, gbs, gbs / PEAK), flush=True)
        del g, env
        torch.cuda.empty_cache()
