"""Write-only HBM stream probes -> gpurun_out/write_peak.json"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch
from formation_gym import _native as nat
lib = nat.load()
nbytes = 2 << 30
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {}
def run(name, variant, chunk, ctas):
    best = None
    for r in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nat.check(lib.fg_write_probe(variant, buf.data_ptr(), nbytes, chunk, ctas, st), "probe"); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r and (best is None or ms < best): best = ms
    res[name] = nbytes / best / 1e6
    print("%-40s %.0f GB/s" % (name, res[name]), flush=True)
run("stg128 streaming, 148x8 ctas", 0, 0, 148 * 8)
run("stg128 streaming, 148x16 ctas", 0, 0, 148 * 16)
for chunk in (2048, 4096, 16384, 65536):
    for ctas in (148 * 2, 148 * 4, 148 * 8):
        if chunk * (ctas // 148) > 200 * 1024: continue
        run("bulk %d B x %d ctas/SM" % (chunk, ctas // 148), 1, chunk, ctas)
        run("bulk evict_first %d B x %d ctas/SM" % (chunk, ctas // 148), 2, chunk, ctas)
# torch memset and copy for reference
for name, fn in (("torch fill_", lambda: buf.fill_(1)), ("torch copy (read+write)", lambda: buf[:nbytes // 2].copy_(buf[nbytes // 2:]))):
    best = None
    for r in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r and (best is None or ms < best): best = ms
    res[name] = nbytes / best / 1e6
    print("%-40s %.0f GB/s" % (name, res[name]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "write_peak.json"), "w"), indent=1)
