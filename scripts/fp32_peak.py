"""Measure the FP32 probes on the GPU box and write gpurun_out/fp32_peak.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-formation_b200"))
import torch  # noqa: E402
from formation_gym import probe  # noqa: E402

res = {}
for v in ("ffma", "ffma2", "fmnmx", "pairmix"):
    for cps in (4, 8):
        r = probe.measure(v, ctas_per_sm=cps)
        res["%s_ctas%d" % (v, cps)] = r
        # instruction issue rate per SMSP per clock at 1965 MHz
        r["warp_inst_per_clk_per_smsp_at_1965MHz"] = r["thread_inst_per_s"] / 32 / (r["sms"] * 4) / 1.965e9
        print(v, cps, "%.3f ms  %.1f TFLOP/s  %.3f warp-inst/clk/SMSP" % (r["ms"], r["tflops"], r["warp_inst_per_clk_per_smsp_at_1965MHz"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fp32_peak.json"), "w"), indent=1)
