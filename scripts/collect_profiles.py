"""Turn the CSV exports of scripts/round_profiles.sh (gpurun_out/TAG_*) into the committed evidence under profiles/:
TAG_<config>_ncu_summary.json (selected raw-page metrics) and TAG_<config>_hot_lines.txt (stall samples per source line).
Usage: python scripts/collect_profiles.py TAG"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_csv_summary  # noqa: E402
tag = sys.argv[1]
short = {"formation_hd_partial_env": "partial", "formation_hd_obs_env": "obstacle", "formation_hd_env": "hd",
         "basic_formation_env": "basic"}
for raw in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag + "_*.raw.csv"))):
    name = os.path.basename(raw)[:-8]
    for k, v in short.items():
        name = name.replace(k, v)
    out = os.path.join(ROOT, "profiles", name + "_ncu_summary.json")
    ncu_csv_summary.main(raw, out)
    src = raw[:-8] + ".source.csv.gz"
    if os.path.isfile(src):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_line_hot.py"), src, "40"],
                             capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", name + "_hot_lines.txt"), "w").write(txt)
    print("wrote", out)
