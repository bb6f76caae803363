"""Exploration: write-only stream with SMALL TMA bulk stores (row pieces of small-N observation rows)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch
from formation_gym import _native as nat
lib = nat.load()
nbytes = 1 << 30
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for chunk in (64, 128, 224, 432, 656, 1024, 2048, 5824, 17488):
    for per_sm in (8, 16, 32):
        best = None
        for r in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); nat.check(lib.fg_write_probe(2, buf.data_ptr(), nbytes, chunk, 148 * per_sm, st), "probe"); b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            if r and (best is None or ms < best): best = ms
        print("bulk evict_first %6d B x %2d issuers/SM: %6.0f GB/s  %.1f Mops/s/SM" %
              (chunk, per_sm, nbytes / best / 1e6, nbytes / chunk / best / 1e3 / 148), flush=True)
