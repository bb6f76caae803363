"""Summarise the raw-page CSV that scripts/gpu_prof.sh exports on the GPU box (ncu -i rep --page raw --csv) into the
small JSON kept under profiles/.  Usage: python scripts/ncu_csv_summary.py in.raw.csv out.json"""
import csv
import json
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_summary import KEYS  # noqa: E402

EXTRA = ["dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_write.sum",
         "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "sm__inst_executed_pipe_lsu.sum",
         "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "smsp__warps_eligible.avg.per_cycle_active",
         "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
         "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]


def main(src, out):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = {"_source": src, "_how": "ncu --set full --clock-control none (see scripts/gpu_prof.sh for --cache-control)"}
    for k in KEYS + EXTRA:
        if k in hdr:
            i = hdr.index(k)
            res[k] = {"unit": units[i], "values": [r[i] for r in data]}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    return res


if __name__ == "__main__":
    r = main(sys.argv[1], sys.argv[2])
    for k, v in r.items():
        if not k.startswith("_"):
            print("%-75s %s %s" % (k[:75], v["values"][0], v["unit"]))
