"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass) -> profiles/sass_digest.txt.
Shows which Blackwell-specific instructions each kernel family uses: UBLKCP (TMA bulk copy, cp.async.bulk),
FFMA2 / FADD2 / FMUL2 (packed fp32 pairs), FMNMX3 (3-input min/max), REDUX (warp reduce), CCTL/PREFETCH (L2 prefetch),
LDGSTS (cp.async), UTC*MMA / LDTM / STTM (tcgen05: none expected -- nothing on this path is a contraction).
Usage: python scripts/sass_digest.py [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gym-formation_b200", "formation_gym", "libformation_gym_b200.so")
WATCH = ["UBLKCP", "UTMASTG", "UTMALDG", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "REDUX", "CCTL", "LDGSTS", "MUFU",
         "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "BAR",
         "SHFL", "VOTE", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern][op] += 1; counts[kern]["_all"] += 1
names = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
fam = collections.OrderedDict()
for k, nm in zip(counts, names):
    nm = re.sub(r"\(fg::[KP]Args<T1>\)|void ", "", nm)
    f = re.sub(r"<.*", "", nm)
    fam.setdefault(f, []).append((nm, counts[k]))
lines = ["SASS digest of %s (cuobjdump -sass; sm_100a).  Counts are static instruction counts per kernel." % os.path.basename(SO),
         "Columns: " + " ".join(WATCH), ""]
for f, ks in fam.items():
    tot = collections.Counter()
    for _, c in ks:
        tot.update(c)
    lines.append("== %s: %d instantiation(s), %d SASS instructions" % (f, len(ks), tot["_all"]))
    lines.append("   " + "  ".join("%s=%d" % (w, tot[w]) for w in WATCH if tot[w]))
lines.append("")
lines.append("-- headline / named instantiations")
for f, ks in fam.items():
    for nm, c in ks:
        if re.search(r"k_hd_warp<float, \(int\)(3|9|27), \(bool\)1, \(int\)[01], \(bool\)1>|k_step<float, \(int\)0, \(bool\)1, \(bool\)1, \(bool\)0, \(int\)[12], \(bool\)1>|k_policy_bfs<float, \(int\)3, \(int\)2>|k_rows_to_host<float2>", nm):
            lines.append("%s: %d instr; " % (nm, c["_all"]) + "  ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w]))
txt = "\n".join(lines) + "\n"
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_digest.txt")
open(dst, "w").write(txt)
print(txt[:3000])
