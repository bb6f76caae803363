"""FP32 pipe probes (diagnostics): measures the FP32 peak of the current GPU with the library's
``fg_fp32_probe`` kernels, timed with CUDA events.  Used by bench.py for the large-N roofline."""
import ctypes as C

import torch

from . import _native as nat

# (variant, flop per lane per counted instruction, counted instructions per iteration and thread)
VARIANTS = {
    "ffma": (0, 2.0, 32),          # scalar FFMA
    "ffma2": (1, 4.0, 32),         # packed fma.rn.f32x2
    "fmnmx": (2, 1.0, 32),         # alu pipe
    "pairmix": (3, 17.0 / 7.0, 56),   # per pair-step: 3 FADD2 (6) + FMUL2 (2) + FFMA2 (4) ... see below
}


def measure(variant="ffma2", iters=4096, ctas_per_sm=8, reps=5, device=None):
    """Returns dict(tflops, warp_inst_per_clk_per_smsp_at_max_clock, ms)."""
    lib = nat.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    v, flop, ninst = VARIANTS[variant]
    ctas = sms * ctas_per_sm
    scratch = torch.zeros(ctas * 256, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    best = None
    for r in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(lib.fg_fp32_probe(v, iters, ctas, scratch.data_ptr(), st), "fg_fp32_probe")
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b)
        if r > 0 and (best is None or ms < best):
            best = ms
    inst = float(ctas) * 256 * iters * ninst                      # thread instructions
    if variant == "pairmix":
        # per pair-step (two pairs): FADD2 x4 (8 flop) + FMUL2 (2) + FFMA2 (4) + FMNMX3 (2 compares) = 16 flop
        flops = float(ctas) * 256 * iters * 8 * 16.0
    else:
        flops = inst * flop
    return {"variant": variant, "ms": best, "tflops": flops / (best * 1e-3) / 1e12,
            "thread_inst_per_s": inst / (best * 1e-3), "sms": sms}


def fp32_peak(device=None):
    """Best of the scalar and packed FMA probes, TFLOP/s."""
    a = measure("ffma", device=device)
    b = measure("ffma2", device=device)
    return max(a["tflops"], b["tflops"]), {"ffma": a, "ffma2": b}
