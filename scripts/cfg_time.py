"""Exploration helper: time the fused step (CUDA-graph replay, actions pre-sampled) for a list of
configurations.  Usage: python scripts/cfg_time.py "hd:243:1024:1" "hd:243:1024:0" ...  (scn:N:E:obs)"""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import sweep  # noqa: E402

for spec in sys.argv[1:]:
    scn, N, E, obs = spec.split(":")
    us, gbs = sweep.run(int(N), int(E), obs == "1", only_step=True, reps=8,
                        scenario={"hd": "formation_hd_env", "basic": "basic_formation_env"}[scn])
    print("%s N=%s E=%s obs=%s: %.1f us/step  %.0f GB/s algorithmic  %.3g agent-steps/s" %
          (scn, N, E, obs, us, gbs, int(N) * int(E) / us * 1e6), flush=True)
