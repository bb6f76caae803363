"""Opcode histogram + hottest SASS lines of an exported ncu source page (scripts/gpu_prof.sh *.source.csv.gz)."""
import csv, gzip, sys, collections, re
f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(gzip.open(f, "rt")))
hdr = rows[1]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot_i = sum(int(r[iI]) for r in data); tot_s = sum(int(r[iN]) for r in data)
ops = collections.Counter(); smp = collections.Counter()
for r in data:
    op = r[iS].strip().split()[0]
    if op.startswith("@"): op = r[iS].strip().split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[iI]); smp[op] += int(r[iN])
print("total warp instructions %d, samples %d, SASS lines %d" % (tot_i, tot_s, len(data)))
for op, n in ops.most_common(28):
    print("  %-10s %6.2f%% instr  %6.2f%% samples" % (op, 100.0 * n / tot_i, 100.0 * smp[op] / max(tot_s, 1)))
print("hottest lines:")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[iN]))[:top]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("  %5.2f%% %-70s %s" % (100.0 * int(r[iN]) / max(tot_s, 1), r[iS].strip()[:70], st))
