"""Profiling helper (not part of the product): steps one batched env a few times so that ncu can
capture the fused step kernel of a chosen configuration.
Usage: python scripts/run_cfg.py SCENARIO N E [steps] [obs=1|0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
import formation_gym  # noqa: E402

scen, N, E = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
obs = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
env = formation_gym.make_batched_env(scen, E, N, 25, write_obs=obs, seed=1)
env.reset()
for _ in range(steps):
    env.sample_actions()
    env.step(env.actions)
torch.cuda.synchronize()
print("ok", scen, N, E, steps, obs)
