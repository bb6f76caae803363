"""Exploration: step time with (a) the same pre-sampled actions every step, (b) fresh random actions every step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import sweep
for N, E in ((9, 131072), (27, 65536), (3, 1048576)):
    for only in (True, False):
        us, gbs = sweep.run(N, E, True, only_step=only, reps=8)
        print("N=%d E=%d %s: %.1f us/step %.0f GB/s  [%s]" % (N, E, "same actions" if only else "fresh actions (+policy kernel)", us, gbs, sweep.LAST), flush=True)
