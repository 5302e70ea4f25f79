"""Host emulation of the CUDA-core kernels — TEST INFRASTRUCTURE ONLY.

The same ``jperceiver_b200/csrc/*.cu`` sources (minus the tcgen05/TMA files) are compiled as plain
C++ with ``-DJPB_HOST_EMU`` (see ``csrc/jpb_common.cuh``): one "thread" per block, blocks run in
sequence.  It exists so kernel *logic* can be checked against the oracle in the GPU-less authoring
container; it is slow, never shipped, and ``jperceiver_b200`` never loads it on its own — tests
install it explicitly through ``jperceiver_b200._lib.use_library``.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "jperceiver_b200", "csrc")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
# files that contain inline PTX / TMA and cannot be emulated
EXCLUDE = ("conv_tc.cu", "tma_util.cu")


def build_emulation(mt: bool | None = None) -> str:
    """``mt=False``: one "thread" per block (fast; kernel arithmetic and indexing).  ``mt=True``: ``-DJPB_HOST_EMU_MT`` — every
    block runs with its real thread count on OS threads with real barriers / shuffles (slow; synchronisation and reductions)."""
    from .torch_ops import install_conv_patch
    install_conv_patch()   # the emulation build has no tcgen05 convolution: library convolution while it is installed
    if mt is None:   # JPB_EMU_MT=1 python -m pytest tests/<file> -m "not gpu": any emulation suite with real block threads
        mt = os.environ.get("JPB_EMU_MT", "0") not in ("", "0")
    os.makedirs(OUT, exist_ok=True)
    srcs = [s for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if os.path.basename(s) not in EXCLUDE]
    h = hashlib.sha256()
    for p in srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "jpb200.h")]:
        h.update(open(p, "rb").read())
    h.update(b"flags-v2")
    stem = "libjpb200_emt_" if mt else "libjpb200_emu_"
    lib = os.path.join(OUT, stem + "%s.so" % h.hexdigest()[:12])
    if os.path.exists(lib):
        return lib
    for old in glob.glob(os.path.join(OUT, stem + "*.so")):
        os.remove(old)
    # -fno-gnu-unique / -Bsymbolic: the two emulation libraries define the same inline state (jpb_common.cuh) and may be loaded into
    # one process; each must bind to its own copy
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-gnu-unique", "-Wl,-Bsymbolic", "-DJPB_HOST_EMU", "-Wno-unused-variable",
           "-Wno-unused-function", "-o", lib]
    if mt:
        cmd[2] = "-std=c++20"
        cmd += ["-DJPB_HOST_EMU_MT", "-pthread"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stdout + r.stderr)
    return lib
