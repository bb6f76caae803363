import os, sys
sys.path.insert(0, "/root/repo/gym-formation_b200")
import torch, formation_gym
for E in (592, 1024, 1184, 2048, 4096):
    env = formation_gym.make_batched_env("formation_hd_env", E, 243, 1000, seed=1)
    env.reset(); env.sample_actions()
    g = env.capture_steps(5, policy=lambda e_: None)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6): g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print("E=%5d  %8.2f us" % (E, best * 1e3), flush=True)
    del g, env; torch.cuda.empty_cache()
