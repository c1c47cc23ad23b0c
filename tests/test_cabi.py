"""The C-ABI boundary without a GPU: libb200av1.so builds, loads, exports every symbol include/b200av1.h
declares, the ctypes binding declares exactly the same set, and every ABI struct has the same size on both sides."""
import ctypes as C
import os

import refs
from dav1d_b200 import _lib, synth


def test_cabi_exports_every_declared_symbol():
    import re
    from dav1d_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(refs.ROOT, "include", "b200av1.h")).read()
    declared = set(re.findall(r"B200_API\s+[^;(]*?\b(b200_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.B200Lib(_lib.LIB_PATH)          # loads; resolves every bound symbol
    for name in declared:
        assert hasattr(lib.dll, name), "library does not export " + name
    assert declared == set(_lib.B200Lib.symbols()), declared ^ set(_lib.B200Lib.symbols())
    assert lib.b200_version() >= 100



def test_abi_struct_sizes_match_binding():
    from dav1d_b200 import build
    build.build()
    lib = _lib.B200Lib(_lib.LIB_PATH)            # raises on any sizeof mismatch
    assert lib.b200_struct_size(9) == C.sizeof(_lib.FrameJob)
    assert C.sizeof(_lib.Av1Filter) == synth.AV1FILTER_DT.itemsize == 1348
    assert synth.MC_BLOCK_DT.itemsize == C.sizeof(_lib.McBlock) and synth.COMP_BLOCK_DT.itemsize == C.sizeof(_lib.CompBlock)


