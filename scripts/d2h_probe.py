"""Diagnostic (not part of the product): how fast can one step's observations reach a pinned host array?
Times fg_obs_to_host modes 0-3 for hd N agents x E envs.  Usage: python scripts/d2h_probe.py [N] [E]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
from formation_gym import _native as nat  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 9
E = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
lib = nat.load()
dev = torch.device("cuda", 0)
rows, row_items, dyn = E * N, 3 * N, N
obs = torch.randn(E, N, 6 * N, device=dev)
done = torch.zeros(E, N, dtype=torch.uint8, device=dev)
host = torch.zeros(E, N, 6 * N).pin_memory()
stage_d = torch.empty(rows, 2 * dyn, device=dev)
stage_h = torch.empty(rows, 2 * dyn).pin_memory()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def run(mode, dst, reps=10, done_t=None):
    def once():
        nat.check(lib.fg_obs_to_host(obs.data_ptr(), dst.data_ptr(), None if done_t is None else done_t.data_ptr(),
                                     stage_d.data_ptr(), E, N, row_items, dyn, 8, mode, st), "fg_obs_to_host")
        torch.cuda.current_stream(dev).synchronize()
    once(); once()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    return (time.perf_counter() - t0) / reps


full_b, dyn_b = obs.numel() * 4, rows * dyn * 8
for mode, name, dst, b in ((0, "full contiguous memcpy", host, full_b), (1, "2-D memcpy, dynamic prefix", host, dyn_b),
                           (2, "zero-copy scatter kernel", host, dyn_b), (3, "pack + contiguous memcpy", stage_h, dyn_b),
                           (4, "zero-copy kernel, whole 64-B lines", host, dyn_b)):
    t = run(mode, dst)
    print("mode %d %-32s %8.3f ms  %7.2f GB/s useful (%d MB)" % (mode, name, t * 1e3, b / t / 1e9, b >> 20), flush=True)
# correctness of mode 1 / 2 (host array must equal the device tensor where written)
host.zero_(); run(2, host, 1)
ok2 = torch.equal(host[:, :, :2 * N], obs.cpu()[:, :, :2 * N]) and float(host[:, :, 2 * N:].abs().sum()) == 0.0
host.copy_(obs.cpu()); host[:, :, :2 * N] = 0; run(4, host, 1)            # mode 4 relies on the static part being in place
ok4 = torch.equal(host, obs.cpu())
host.zero_(); run(1, host, 1)
ok1 = torch.equal(host[:, :, :2 * N], obs.cpu()[:, :, :2 * N]) and float(host[:, :, 2 * N:].abs().sum()) == 0.0
done[::7] = 1
host.zero_(); run(2, host, 1, done)
oc = obs.cpu()
ok2d = torch.equal(host[::7], oc[::7]) and torch.equal(host[1::7, :, :2 * N], oc[1::7, :, :2 * N]) and \
    float(host[1::7, :, 2 * N:].abs().sum()) == 0.0
print("correct: mode1", ok1, "mode2", ok2, "mode2+done", ok2d, "mode4", ok4)
# host-side scatter of the packed staging array (torch CPU, multi-threaded strided copy)
hv = host.view(rows, 6 * N)
t0 = time.perf_counter()
for _ in range(5):
    hv[:, :2 * N].copy_(stage_h)
t = (time.perf_counter() - t0) / 5
print("host scatter of packed rows (torch, %d threads): %.3f ms" % (torch.get_num_threads(), t * 1e3))
# all envs done (mode 2 copies whole rows)
done.fill_(1)
t = run(2, host, 5, done)
print("mode 2, every env done (whole rows): %.3f ms  %.2f GB/s" % (t * 1e3, full_b / t / 1e9))

# ---- 2-D copies split over K streams (K copy-engine queues): does the small-row DMA scale beyond one engine?
for K in (1, 2, 3, 4, 6, 8):
    streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
    per = E // K

    def once():
        for k, s in enumerate(streams):
            e0 = k * per
            ne = per if k < K - 1 else E - e0
            off = e0 * N * 6 * N * 4
            nat.check(lib.fg_obs_to_host(obs.data_ptr() + off, host.data_ptr() + off, None, None, ne, N, row_items, dyn, 8, 1,
                                         C.c_void_p(s.cuda_stream)), "fg_obs_to_host")
        for s in streams:
            s.synchronize()
    once(); once()
    t0 = time.perf_counter()
    for _ in range(10):
        once()
    t = (time.perf_counter() - t0) / 10
    print("mode 1 over %d streams: %.3f ms  %.2f GB/s useful" % (K, t * 1e3, dyn_b / t / 1e9), flush=True)
