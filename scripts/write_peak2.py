"""Write probe mimicking the N=27 warp kernel's store pattern: 17488-byte bulk stores (16-byte, not
128-byte aligned), few in flight per SM."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch
from formation_gym import _native as nat
lib = nat.load()
nbytes = 1 << 30
buf = torch.empty(nbytes + 4096, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(name, variant, chunk, ctas, off=0):
    best = None
    for r in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nat.check(lib.fg_write_probe(variant, buf.data_ptr() + off, nbytes, chunk, ctas, st), "probe"); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r and (best is None or ms < best): best = ms
    print("%-60s %.0f GB/s" % (name, (nbytes // chunk * chunk) / best / 1e6), flush=True)
for chunk in (17488, 17408, 5824, 5888):
    for cps in (3, 6, 12):
        run("bulk evict_first %d B x %d ctas/SM (5 in flight each)" % (chunk, cps), 2, chunk, 148 * cps)
        run("bulk evict_first %d B x %d ctas/SM, base+16" % (chunk, cps), 2, chunk, 148 * cps, 16)
