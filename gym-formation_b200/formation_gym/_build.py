"""Build the sm_100a shared library in-tree with nvcc (no JIT cache: the .so travels with the repo).

    python -m formation_gym._build          # or __graft_entry__.build()
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG_DIR), "csrc")
REPO = os.path.dirname(os.path.dirname(PKG_DIR))
LIB_NAME = "libformation_gym_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# (-split-compile would run ptxas per kernel in parallel, but it changes register allocation: the N = 9 warp
# kernel spilled 120 B instead of 32 B and ran 55 % slower.  The translation units are compiled in parallel
# instead, with unchanged code generation.)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(REPO, "include", "formation_gym_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> formation_gym/libformation_gym_b200.so for sm_100a (one nvcc process per translation
    unit, in parallel; then one link step)."""
    # A/B builds (profiling only): FG_BUILD_DEFINES="-DFG_X=1 ..." adds preprocessor defines and FG_BUILD_OUT names
    # another output file, which `FG_B200_LIB=<that file>` then loads instead of the product library
    extra = os.environ.get("FG_BUILD_DEFINES", "").split()
    out_path = os.environ.get("FG_BUILD_OUT", LIB_PATH)
    if not extra and out_path == LIB_PATH and not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB_NAME)
    objdir = os.path.join(REPO, "build", "obj" if out_path == LIB_PATH else "obj_" + os.path.basename(out_path))
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-I", os.path.join(REPO, "include"), "-c", src, "-o", obj]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    objs = []
    for cmd, obj, proc in jobs:
        _, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), err))
        if verbose:
            sys.stderr.write(err)
        objs.append(obj)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
            "-o", out_path + ".tmp"] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(link), res.stderr))
    os.replace(out_path + ".tmp", out_path)
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
