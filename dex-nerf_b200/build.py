"""Build libdexnerf.so (the C-ABI library of include/dexnerf.h) in-tree with nvcc for sm_100a.

    python dex-nerf_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Objects are cached by source mtime; the shared library lands
in dex-nerf_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdexnerf.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "dexnerf.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, obj, verbose):
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return r.stderr


def build_library(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment builds (tools/): extra -D flags, objects and library under their own names, e.g.
    python dex-nerf_b200/build.py --define DEXNERF_WIDE_EPI=0 --out libdexnerf_epi0.so; select with DEXNERF_LIB."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_m = _deps_mtime()
    jobs, objs = [], []
    lib = os.path.join(LIBDIR, out) if out else LIB
    tag = ("." + os.path.splitext(out)[0]) if out else ""
    global FLAGS
    base_flags = FLAGS
    FLAGS = list(FLAGS) + ["-D" + d for d in defines]
    for f in srcs:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f[:-3] + tag + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append((src, obj))
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(lambda a: _compile(a[0], a[1], verbose), jobs))
    FLAGS = base_flags
    if jobs or not os.path.exists(lib):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("\n".join(logs))
    return lib


if __name__ == "__main__":
    a = sys.argv
    print(build_library(force="--force" in a, verbose="-v" in a,
                        defines=[a[i + 1] for i, x in enumerate(a) if x == "--define"],
                        out=a[a.index("--out") + 1] if "--out" in a else None))
