"""Build libmdseg_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python mul-datasets-semantic-segmentation_b200/build.py [--force] [--verbose]

Every csrc/*.cu is compiled to an object (in parallel, only when stale) and
linked into `libmdseg_b200.so` next to this file.  The shared object has a pure
C ABI (include/mdseg.h); the Python host side loads it with ctypes.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmdseg_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-I", INCLUDE]
CFLAGS += os.environ.get("MDSEG_CFLAGS", "").split()  # experiment switches (-DMDSEG_...=k); empty in normal builds


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, (r.stdout + r.stderr)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    dep_m = _deps_mtime()
    todo = []
    for src in _sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        sm = max(os.path.getmtime(os.path.join(CSRC, src)), dep_m)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, rc, out in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
