"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- recipe that stages the UNMODIFIED reference for the GPU box.

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference exists

The reference (jc-bao/gym-formation) is pure Python, so "building" it is a byte-for-byte copy of the files
on the step path -- formation_gym/{__init__,core,environment,scenario}.py and formation_gym/envs/*.py --
from /root/reference into the git-ignored directory ``oracle/_ref/`` (listed in .gitignore, NOT in
.gpurunignore: it travels to the GPU box with the snapshot exactly like the built .so does, and never
enters the history).  ``oracle/ref_harness.py`` then runs those files unchanged behind its in-process stubs
for the missing ``imp`` / ``gym`` / ``multiagent`` modules, which makes the CPU arm of bench.py
(``--impl reference`` and ``cpu_baseline``) the reference's own code (``kind: "reference"``) instead of the
loop port.  A MANIFEST.json with the sha256 of every staged file is written next to them; the harness
refuses a tree whose hashes do not match it.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
TOP = ("__init__.py", "core.py", "environment.py", "scenario.py")


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged_ok(dst=DST):
    """True when oracle/_ref holds a complete staged tree whose files match its manifest."""
    man = os.path.join(dst, "MANIFEST.json")
    if not os.path.isfile(man):
        return False
    try:
        with open(man) as f:
            m = json.load(f)
        return all(sha256(os.path.join(dst, rel)) == h for rel, h in m["files"].items()) and len(m["files"]) >= 6
    except Exception:
        return False


def make(src=SRC, dst=DST, quiet=False):
    pkg = os.path.join(src, "formation_gym")
    if not os.path.isfile(os.path.join(pkg, "core.py")):
        raise RuntimeError("reference tree not found at %s" % src)
    rels = ["formation_gym/" + f for f in TOP]
    rels += sorted("formation_gym/envs/" + os.path.basename(p) for p in glob.glob(os.path.join(pkg, "envs", "*.py")))
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    files = {}
    for rel in rels:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        files[rel] = sha256(out)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "what": "unmodified files of jc-bao/gym-formation on the MPE step path",
                   "files": files}, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files into %s" % (len(files), dst))
    return dst


if __name__ == "__main__":
    make()
    sys.exit(0 if staged_ok() else 1)
