"""Build ``libjpb200.so`` (all of ``csrc/*.cu``) for sm_100a with nvcc, in-tree.

``python -m jperceiver_b200.build`` — also called by ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU.  The library links only cudart (+ libcuda for the TMA descriptor
encoder); no torch types cross the boundary (see ``include/jpb200.h``).
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libjpb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stamp(sources):
    h = hashlib.sha256()
    for p in sorted(sources + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "jpb200.h")]):
        h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    stamp_file = LIB + ".stamp"
    stamp = _stamp(sources)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    log = []
    for src in sources:
        obj = src[:-3] + ".o"
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    open(stamp_file, "w").write(stamp)
    open(os.path.join(CSRC, "ptxas.log"), "w").write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
