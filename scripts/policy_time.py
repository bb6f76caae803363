"""Exploration helper: time fg_policy_bfs alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch, formation_gym
for N, E in ((9, 131072), (27, 65536), (3, 1048576), (243, 1024), (81, 8192)):
    env = formation_gym.make_batched_env("formation_hd_env", E, N, 25, seed=1, write_obs=False)
    env.reset()
    for _ in range(3): env.bfs_actions(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): env.bfs_actions(3)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    print("bfs policy N=%d E=%d: %.1f us  (%.3g agent-actions/s, %.0f GB/s of state traffic)" % (N, E, us, N * E / us * 1e6, E * (24 * N + 8) / us / 1e3), flush=True)
