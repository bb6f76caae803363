"""Write-only HBM stream of TMA bulk stores (fg_write_probe variant 2: evict_first, as the step kernels) over a
footprint larger than L2, for the piece sizes the warp kernels emit per span -- the ceiling the streaming regime
(>= 256 K envs at N = 9) can be compared with.  Usage: python scripts/write_stream.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch
from formation_gym import _native as nat
lib = nat.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for gb in (2.25,):
    nbytes = int(gb * (1 << 30))
    buf = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    for variant in (0, 2, 4):
        for chunk in (2160, 5824, 17488) if variant else (16,):
            for per_sm in (4, 8, 16, 24):
                best = None
                for r in range(4):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); nat.check(lib.fg_write_probe(variant, buf.data_ptr(), nbytes, chunk, 148 * per_sm, st), "probe"); b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b)
                    if r and (best is None or ms < best): best = ms
                print("%.2f GiB, variant %d, piece %6d B, %d CTAs/SM: %6.0f GB/s" % (gb, variant, chunk, per_sm, nbytes / best / 1e6), flush=True)
    del buf
