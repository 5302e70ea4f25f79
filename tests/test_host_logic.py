"""Host-side logic that needs no GPU: the C-ABI library builds, loads and exports every symbol include/jpb200.h declares
(no compute calls), the implicit-GEMM K-chunk tables and their channel-block-major re-ordering are consistent with the weight
layout, the split-K rule, and the missing-library / CPU-tensor error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from jperceiver_b200 import _lib, conv as JC


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "jpb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jpb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from jperceiver_b200 import build as B
    lib = B.build()
    assert os.path.exists(lib)
    h = ctypes.CDLL(lib)
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(h, name), "include/jpb200.h declares %s but libjpb200.so does not export it" % name
    # and the Python binding knows every one of them (argtypes declared next to the header)
    bound = set(_lib.exported_symbols())
    missing = [n for n in declared if n not in bound]
    assert not missing, missing


def test_chunk_table_matches_weight_layout():
    """One row per 16-byte chunk of K in weight order [tap][source][channel]; rows are (source | tapsrc<<8, dy<<16|dx, channel
    offset, valid bytes); padded to whole K blocks of 8 chunks with -1 rows."""
    t = JC.chunk_table([8, 5], 3, 3, "cpu")
    assert t.shape[1] == 4 and t.shape[0] % 8 == 0
    live = t[t[:, 0] >= 0]
    assert live.shape[0] == 9 * (2 + 2)                      # 8 channels = 2 chunks, 5 channels = 2 chunks (16 + 4 bytes)
    # K position of a chunk in the packed weight: tap * (8 + pad4(5)) + source offset + channel offset
    pos = []
    for x, yx, coff, nbytes in live.tolist():
        si, tapsrc = x & 0xff, x >> 8
        tap = tapsrc // 2
        assert tapsrc % 2 == si and (yx >> 16) == tap // 3 and (yx & 0xffff) == tap % 3
        assert nbytes == (16 if (si == 0 or coff == 0) else 4)
        pos.append(tap * 16 + (0 if si == 0 else 8) + coff)
    assert pos == sorted(pos) and len(set(pos)) == len(pos)


@pytest.mark.parametrize("srcs,k", [([64], 3), ([256, 256, 1], 3), ([128], 1), ([4], 7)])
def test_kblock_reordering_is_a_permutation(srcs, k):
    """The channel-block-major order must visit every K block exactly once and carry the matching weight column."""
    base = JC.chunk_table(srcs, k, k, "cpu")
    tab, kcol = JC.ordered_table(srcs, k, k, "cpu")
    if kcol is None:
        assert k == 1 and tab is base
        return
    nkb = base.shape[0] // 8
    cols = (kcol // 32).tolist()
    assert sorted(cols) == list(range(nkb))
    for i, c in enumerate(cols):
        assert torch.equal(tab[i * 8:(i + 1) * 8], base[c * 8:(c + 1) * 8])
    # taps of one channel block are consecutive: the first chunk's channel offset never decreases within a source
    first = tab.view(-1, 8, 4)[:, 0]
    live = first[first[:, 0] >= 0]
    key = [(int(x) & 0xff, int(c) // 32) for x, _, c, _ in live.tolist()]
    assert key == sorted(key)


def test_split_k_rule():
    assert JC._ksplit(81920, 256, 72) == 1                   # 640 tiles: fills the machine
    ks = JC._ksplit(1280, 512, 144)                          # layer4: 20 tiles
    assert ks > 1 and 20 * ks <= 296 and 144 // ks >= 4
    assert JC._ksplit(1280, 512, 8) == 1                     # too few K blocks to split
    slots, JC.KSPLIT_SLOTS = JC.KSPLIT_SLOTS, 0
    try:
        assert JC._ksplit(1280, 512, 144) == 1
    finally:
        JC.KSPLIT_SLOTS = slots


def test_product_path_refuses_cpu_tensors_and_missing_library():
    _lib._handle, _lib._emulated = None, False
    with pytest.raises(_lib.JpbError):
        _lib.ptr(torch.zeros(4))                             # CPU tensor, no emulation installed: no CPU fallback
    real = _lib.LIB_PATH if hasattr(_lib, "LIB_PATH") else None
    from jperceiver_b200 import netops as ops
    with pytest.raises(_lib.JpbError):
        ops.conv2d(torch.zeros(1, 4, 8, 8), torch.zeros(4, 4, 3, 3), pad=1)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every argument struct of include/jpb200.h against its ctypes mirror in jperceiver_b200/_lib.py: same size, same number of
    members, same offset and size per member (compiled with gcc from the header itself).  Guards the structs that no CPU test can
    exercise (the tcgen05 convolution arguments)."""
    import ctypes as C
    import re
    import subprocess
    hdr = os.path.join(ROOT, "include", "jpb200.h")
    text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        names = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.search(r"(\w+)\s*(\[[^\]]*\])*\s*$", first).group(1))
            names += [re.search(r"(\w+)", r).group(1) for r in rest]
        structs[m.group(3)] = names
    mirror = {"JpbPhotoArgs": _lib.PhotoArgs, "JpbPhotoGrad": _lib.PhotoGrad, "JpbPyramid": _lib.Pyramid, "JpbScaleLabelArgs": _lib.ScaleLabelArgs,
              "JpbScaleLossArgs": _lib.ScaleLossArgs, "JpbBevArgs": _lib.BevArgs, "JpbConvArgs": _lib.ConvArgs, "JpbConvWgradArgs": _lib.ConvWgradArgs,
              "JpbWeightT": _lib.WeightT, "JpbDepthEvalArgs": _lib.DepthEvalArgs, "JpbResizeArgs": _lib.ResizeArgs, "JpbJitterArgs": _lib.JitterArgs,
              "JpbAdamArgs": _lib.AdamArgs}
    assert set(mirror) == set(structs), (sorted(set(structs) ^ set(mirror)))
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % hdr, "int main(void) {"]
    for name, fields in structs.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fields:
            src.append('printf("%s.%s %%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s*)0)->%s));' % (name, f, name, f, name, f))
    src += ["return 0;", "}"]
    cfile, exe = str(tmp_path / "layout.c"), str(tmp_path / "layout")
    open(cfile, "w").write("\n".join(src))
    subprocess.run(["gcc", "-o", exe, cfile], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    layout = {l.split()[0]: tuple(int(v) for v in l.split()[1:]) for l in out if l.strip()}
    for name, cls in mirror.items():
        assert C.sizeof(cls) == layout[name][0], (name, C.sizeof(cls), layout[name])
        assert len(cls._fields_) == len(structs[name]), (name, [f[0] for f in cls._fields_], structs[name])
        for (pyname, *_), cname in zip(cls._fields_, structs[name]):
            d = getattr(cls, pyname)
            assert (d.offset, d.size) == layout["%s.%s" % (name, cname)], (name, pyname, cname, d.offset, d.size, layout["%s.%s" % (name, cname)])
