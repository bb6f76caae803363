"""Exploration: hd observation rows written with plain 8-byte streaming stores from registers (fg_write_probe variant 3)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch
from formation_gym import _native as nat
lib = nat.load()
nbytes = 1 << 30
buf = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for off in (0, 8):
    for N in (9, 27, 81, 243):
        for per_sm in (4, 8, 16):
            best = None
            for r in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); nat.check(lib.fg_write_probe(3, buf.data_ptr() + off, nbytes, N, 148 * per_sm, st), "probe"); b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b)
                if r and (best is None or ms < best): best = ms
            wrote = (nbytes // (24 * N * N)) * 24 * N * N
            print("rows N=%3d (row %5d B) offset %d, %2d x 4 warps/SM: %6.0f GB/s" % (N, 24 * N, off, per_sm, wrote / best / 1e6), flush=True)
