"""Host-side mirror of dav1d's Dav1dDSPContext function-pointer surface (reference
src/internal.h:62-70) over the Level-1 tables of the C ABI.

Each member is the *same C function pointer* the reference-side binding would install
(b200_<family>_dsp_init_{8,16}bpc), wrapped so it can be called with
numpy buffers the way tests/checkasm/*.c call the reference: `dst` is the address of the
top-left pixel (a numpy view), strides are in BYTES and may be negative, the itx callee
zeroes the coefficient block. Every call ends in a CUDA kernel launch; there is no CPU path.
"""
import ctypes as C
import numpy as np

from . import levels as L
from ._lib import get_lib, ITXFM_FN_8, ITXFM_FN_16


def _addr(a):
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class InvTxfmDSPContext:
    """Dav1dInvTxfmDSPContext (reference src/itx.h:70-72): itxfm_add[N_RECT_TX_SIZES][N_TX_TYPES_PLUS_LL]."""

    def __init__(self, bpc, lib=None):
        assert bpc in (8, 10, 12)
        self.bpc = bpc
        self.bitdepth_max = (1 << bpc) - 1
        self.lib = lib or get_lib()
        n = L.N_RECT_TX_SIZES * L.N_TX_TYPES_PLUS_LL
        self._tbl = (C.c_void_p * n)()
        if bpc == 8:
            self.lib.b200_itx_dsp_init_8bpc(self._tbl, bpc)
        else:
            self.lib.b200_itx_dsp_init_16bpc(self._tbl, bpc)
        self.itxfm_add = [[self._wrap(self._tbl[tx * L.N_TX_TYPES_PLUS_LL + tp])
                           for tp in range(L.N_TX_TYPES_PLUS_LL)] for tx in range(L.N_RECT_TX_SIZES)]

    def _wrap(self, ptr):
        if not ptr:
            return None
        if self.bpc == 8:
            fn = ITXFM_FN_8(ptr)
            return lambda dst, stride, coeff, eob: fn(_addr(dst), stride, _addr(coeff), eob)
        fn = ITXFM_FN_16(ptr)
        bdmax = self.bitdepth_max
        return lambda dst, stride, coeff, eob: fn(_addr(dst), stride, _addr(coeff), eob, bdmax)
