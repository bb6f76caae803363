"""Profiling helper: a few launches of the device controller (fg_policy_bfs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch, formation_gym
N, E = int(sys.argv[1]), int(sys.argv[2])
env = formation_gym.make_batched_env("formation_hd_env", E, N, 25, seed=1, write_obs=False)
env.reset()
for _ in range(5): env.bfs_actions(3)
torch.cuda.synchronize()
