"""Per-SOURCE-LINE stall samples of a kernel: joins the SASS page exported by scripts/gpu_prof.sh (*.source.csv.gz:
one row per SASS instruction with sample counts) with `nvdisasm -g` line info of the same kernel in the locally built
library (same build -> same SASS, joined by instruction order).
Usage: python scripts/ncu_line_hot.py X.source.csv.gz [top] [--so path]"""
import collections, csv, gzip, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
so = os.path.join(ROOT, "gym-formation_b200", "formation_gym", "libformation_gym_b200.so")
rows = list(csv.reader(gzip.open(f, "rt")))
kname = rows[0][1]
hdr, data = rows[1], rows[2:]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
cache = os.path.join(tempfile.gettempdir(), "fg_sass_%d" % int(os.path.getmtime(so)))
os.makedirs(cache, exist_ok=True)
if not os.listdir(cache):
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=cache, stdout=subprocess.DEVNULL)
    for cb in os.listdir(cache):
        if cb.endswith(".cubin"):
            with open(os.path.join(cache, cb[:-6] + ".sass"), "w") as out:
                subprocess.call(["nvdisasm", "-g", cb], cwd=cache, stdout=out, stderr=subprocess.DEVNULL)
norm = lambda s: re.sub(r"\(int\)|\(bool\)|\s", "", s).replace("void", "")   # noqa: E731
want = norm(kname.split("(fg::KArgs")[0].split("(fg::PArgs")[0])
lines = None
for sf in os.listdir(cache):
    if not sf.endswith(".sass"):
        continue
    txt = open(os.path.join(cache, sf)).read().split("\n.text.")
    for blk in txt[1:]:
        mangled = blk.split(":", 1)[0]
        dem = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip()
        d = norm(dem.split("(fg::KArgs")[0].split("(fg::PArgs")[0]).replace("true", "1").replace("false", "0")
        if d == want:
            cur, lines = ("?", 0), []
            for ln in blk.split("\n"):
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                    lines.append(cur)
            break
    if lines is not None:
        break
if lines is None:
    sys.exit("kernel not found in the local build: " + kname)
if len(lines) != len(data):
    print("WARNING: SASS length differs (local %d vs profiled %d): the library was rebuilt since the capture" % (len(lines), len(data)))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for (fl, ln), r in zip(lines, data):
    a = agg[(fl, ln)]
    a[0] += int(r[iN]); a[1] += int(r[iI])
    for i in stall_cols:
        a[2][hdr[i][6:]] += int(r[i])
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
src = {}
print(kname); print("samples %d  warp instructions %d" % (tot_s, tot_i))
for (fl, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if fl not in src:
        p = os.path.join(ROOT, "gym-formation_b200", "csrc", fl)
        src[fl] = open(p).read().split("\n") if os.path.isfile(p) else []
    text = src[fl][ln - 1].strip()[:80] if 0 < ln <= len(src[fl]) else ""
    st = ", ".join("%s %d" % (k, v) for k, v in a[2].most_common(3))
    print("%5.2f%% smp %5.2f%% ins  %s:%d  %-80s [%s]" % (100.0 * a[0] / max(tot_s, 1), 100.0 * a[1] / max(tot_i, 1), fl, ln, text, st))
