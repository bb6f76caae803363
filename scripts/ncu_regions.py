"""Group the SASS of an ncu capture into regions of similar execution count (hot loops).
Usage: python scripts/ncu_regions.py rep [min_total]"""
import csv, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); st = hdr.index("# Samples")
cnts = [int(r[ie]) for r in data]; tot = sum(cnts); smp = [int(r[st]) for r in data]; stot = sum(smp)
print("total warp instructions %.2fM, samples %d" % (tot / 1e6, stot))
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(cnts[j + 1] - cnts[i]) <= 0.15 * max(cnts[i], 1): j += 1
    s = sum(cnts[i:j + 1])
    if s > thresh * tot:
        ops = {}
        for r in data[i:j + 1]:
            op = r[1].split()[0] if not r[1].split()[0].startswith("@") else r[1].split()[1]
            op = op.split(".")[0]; ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
        print("instr %4d-%4d n=%3d count/inst~%8d total=%6.2fM (%4.1f%%) samples=%5d (%4.1f%%) %s" % (
            i, j, j - i + 1, cnts[i], s / 1e6, 100 * s / tot, sum(smp[i:j + 1]), 100 * sum(smp[i:j + 1]) / max(stot, 1), top))
    i = j + 1
