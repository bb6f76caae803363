"""Step + device controller: the step kernel alone (pre-sampled actions), controller kernel + step kernel per step, and
the controller compiled into the step kernel (fg_step_policy), each from a CUDA graph of 5 steps.
Usage: python scripts/policy_fused_time.py ["N E n" ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
import formation_gym  # noqa: E402


def timed(g):
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    return best * 1e3


cfgs = sys.argv[1:] or ["9 131072 3", "27 65536 3", "3 1048576 3", "4 524288 2", "8 131072 2", "16 65536 4", "25 65536 5"]
for cfg in cfgs:
    N, E, n = (int(x) for x in cfg.split())
    out = []
    for mode in ("step", "two", "fused"):
        env = formation_gym.make_batched_env("formation_hd_env", E, N, 25, seed=1)
        env.reset()
        if mode == "step":
            env.bfs_actions(n)
            g = env.capture_steps(5, policy=lambda e_: None)
        elif mode == "two":
            g = env.capture_steps(5, policy=lambda e_: e_.bfs_actions(n))
        else:
            g = env.capture_steps(5, fused_bfs=n)
        out.append(timed(g))
        nb = env.bytes_per_env_step() * E
        del g, env
        torch.cuda.empty_cache()
    fr = [nb / (t * 1e-6) / 1e9 / 6547.8 for t in out]
    print("N=%3d E=%8d n=%d  step %7.2f us (%.3f)  controller kernel + step %7.2f us (%.3f)  fused %7.2f us (%.3f)"
          % (N, E, n, out[0], fr[0], out[1], fr[1], out[2], fr[2]), flush=True)
