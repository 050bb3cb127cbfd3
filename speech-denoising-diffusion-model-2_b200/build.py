"""Builds csrc/*.cu into csrc/libsddm_b200.so for sm_100a (in-tree, so the .so travels with the repo snapshot).

    python -m sddm_b200.build            # or  __graft_entry__.build()

nvcc cross-compiles without a GPU.  Objects go to <repo>/build/ (git-ignored); only sources that changed
since the last build are recompiled.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
REPO = os.path.dirname(HERE)
OBJDIR = os.path.join(REPO, "build", "sddm_b200")
LIB = os.path.join(CSRC, "libsddm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-DSDDM_BUILD",
]
# per-file extras: the tensor-core kernel computes Swish with ex2.approx / rcp.approx and flushes denormals (its operands
# are rounded to bf16 anyway); every other translation unit keeps IEEE arithmetic (bit-exact posterior update, fp32 parity mode)
EXTRA_FLAGS = {"conv_tc.cu": ["-use_fast_math"]}
# debug builds: SDDM_NVCC_EXTRA="-DSDDM_ROW_TRACE=1" python -m sddm_b200.build  (per-role wait tracing of conv_row.cu, tools/prof_ops.py --trace)
NVCC_FLAGS += [f for f in os.environ.get("SDDM_NVCC_EXTRA", "").split() if f]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the sddm_b200 CUDA extension cannot be built")
    return exe


def _stamp(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(sorted(EXTRA_FLAGS.items())).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sources()
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(REPO, "include", "sddm_b200.h"))
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp([src] + headers)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            return obj, False
        cmd = [nvcc] + NVCC_FLAGS + EXTRA_FLAGS.get(os.path.basename(src), []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return obj, True

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
