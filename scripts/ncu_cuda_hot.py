"""Hottest CUDA source lines of an exported ncu source page (scripts/gpu_prof.sh *.cuda.csv.gz)."""
import csv, gzip, sys
f = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(gzip.open(f, "rt")))
# find header row
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iL = hdr.index("#") if "#" in hdr else None
def num(x):
    try: return int(float(x))
    except Exception: return 0
tot_s = sum(num(r[iN]) for r in data); tot_i = sum(num(r[iI]) for r in data)
print("samples %d, instructions %d, lines %d" % (tot_s, tot_i, len(data)))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -num(r[iN]))[:top]:
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print("%5.2f%% smp %5.2f%% ins | %s | %-95s %s" % (100.0 * num(r[iN]) / max(tot_s, 1), 100.0 * num(r[iI]) / max(tot_i, 1),
                                                 r[iL] if iL is not None else "", r[iS].strip()[:95], st))
