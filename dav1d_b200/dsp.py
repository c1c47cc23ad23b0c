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


# ------------------------------------------------------------------------------------------
# Function-pointer prototypes of Dav1dMCDSPContext members (reference src/mc.h:38-114), as
# (8 bpc argument list); the 16 bpc variants append `int bitdepth_max` (HIGHBD_DECL_SUFFIX)
# except blend*, emu_edge which carry no bit-depth argument.
_P, _S, _I, _L = C.c_void_p, C.c_ssize_t, C.c_int, C.c_ssize_t
MC_PROTOS = {
    "mc": ([_P, _S, _P, _S, _I, _I, _I, _I], True),
    "mc_scaled": ([_P, _S, _P, _S, _I, _I, _I, _I, _I, _I], True),
    "mct": ([_P, _P, _S, _I, _I, _I, _I], True),
    "mct_scaled": ([_P, _P, _S, _I, _I, _I, _I, _I, _I], True),
    "avg": ([_P, _S, _P, _P, _I, _I], True),
    "w_avg": ([_P, _S, _P, _P, _I, _I, _I], True),
    "mask": ([_P, _S, _P, _P, _I, _I, _P], True),
    "w_mask": ([_P, _S, _P, _P, _I, _I, _P, _I], True),
    "blend": ([_P, _S, _P, _I, _I, _P], False),
    "blend_v": ([_P, _S, _P, _I, _I], False),
    "blend_h": ([_P, _S, _P, _I, _I], False),
    "warp8x8": ([_P, _S, _P, _S, _P, _I, _I], True),
    "warp8x8t": ([_P, _S, _P, _S, _P, _I, _I], True),
    "emu_edge": ([_L, _L, _L, _L, _L, _L, _P, _S, _P, _S], False),
    "resize": ([_P, _S, _P, _S, _I, _I, _I, _I, _I], True),
}
# member order and array lengths of the struct (reference src/mc.h:146-162)
MC_LAYOUT = [("mc", 10), ("mc_scaled", 10), ("mct", 10), ("mct_scaled", 10), ("avg", 1), ("w_avg", 1),
             ("mask", 1), ("w_mask", 3), ("blend", 1), ("blend_v", 1), ("blend_h", 1), ("warp8x8", 1),
             ("warp8x8t", 1), ("emu_edge", 1), ("resize", 1)]


def wrap_dsp_table(tbl, layout, protos, hbd, bitdepth_max):
    """Turn a flat table of C function pointers into attributes of callables taking numpy buffers /
    ints in dav1d's argument order (bitdepth_max appended automatically for 16 bpc)."""
    out, i = {}, 0
    for name, n in layout:
        args, has_bd = protos[name]
        ft = C.CFUNCTYPE(None, *(args + ([_I] if (hbd and has_bd) else [])))
        fns = []
        for _ in range(n):
            p = tbl[i]; i += 1
            if not p:
                fns.append(None)
                continue
            f = ft(p)

            def call(*a, _f=f, _bd=(hbd and has_bd)):
                a = [_addr(x) if isinstance(x, np.ndarray) else x for x in a]
                if _bd:
                    a.append(bitdepth_max)
                return _f(*a)
            fns.append(call)
        out[name] = fns if n > 1 else fns[0]
    return out


class MCDSPContext:
    """Dav1dMCDSPContext (reference src/mc.h:146-162): mc[10] mc_scaled[10] mct[10] mct_scaled[10] avg
    w_avg mask w_mask[3] blend blend_v blend_h warp8x8 warp8x8t emu_edge resize."""

    def __init__(self, bpc, lib=None):
        assert bpc in (8, 10, 12)
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 53)()
        (self.lib.b200_mc_dsp_init_8bpc if bpc == 8 else self.lib.b200_mc_dsp_init_16bpc)(self._tbl)
        for k, v in wrap_dsp_table(self._tbl, MC_LAYOUT, MC_PROTOS, bpc > 8, self.bitdepth_max).items():
            setattr(self, k, v)


LF_PROTO = [_P, _S, _P, _P, _S, _P, _I]   # decl_loopfilter_sb_fn (reference src/loopfilter.h:39-43)


class LoopFilterDSPContext:
    """Dav1dLoopFilterDSPContext (reference src/loopfilter.h:45-53): loop_filter_sb[plane_class][dir]."""

    def __init__(self, bpc, lib=None):
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 4)()
        (self.lib.b200_loop_filter_dsp_init_8bpc if bpc == 8 else self.lib.b200_loop_filter_dsp_init_16bpc)(self._tbl)
        t = wrap_dsp_table(self._tbl, [("f", 4)], {"f": (LF_PROTO, True)}, bpc > 8, self.bitdepth_max)["f"]
        self.loop_filter_sb = [[t[0], t[1]], [t[2], t[3]]]


class CdefDSPContext:
    """Dav1dCdefDSPContext (reference src/cdef.h:64-67): dir, fb[3] (8x8 / 4x8 / 4x4).
    dir(img, stride) -> (dir, var); fb[i](dst, stride, left, top, bottom, pri, sec, dir, damping, edges)."""

    def __init__(self, bpc, lib=None):
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 4)()
        (self.lib.b200_cdef_dsp_init_8bpc if bpc == 8 else self.lib.b200_cdef_dsp_init_16bpc)(self._tbl)
        hbd = bpc > 8
        dir_ft = C.CFUNCTYPE(_I, _P, _S, C.POINTER(C.c_uint), *([_I] if hbd else []))
        fb_ft = C.CFUNCTYPE(None, _P, _S, _P, _P, _P, _I, _I, _I, _I, _I, *([_I] if hbd else []))
        self._dir = dir_ft(self._tbl[0])
        self._fb = [fb_ft(self._tbl[1 + i]) for i in range(3)]
        bd = [self.bitdepth_max] if hbd else []

        def dir_(img, stride):
            var = C.c_uint(0)
            d = self._dir(_addr(img), stride, C.byref(var), *bd)
            return d, var.value
        self.dir = dir_
        self.fb = [(lambda dst, st, left, top, bot, pri, sec, d, damp, edges, _f=f:
                    _f(_addr(dst), st, _addr(left), _addr(top), _addr(bot), pri, sec, d, damp, edges, *bd))
                   for f in self._fb]


LR_PROTO = [_P, _S, _P, _P, _I, _I, _P, _I]   # decl_lr_filter_fn (reference src/looprestoration.h:64-70)


class LoopRestorationDSPContext:
    """Dav1dLoopRestorationDSPContext (reference src/looprestoration.h:72-75): wiener[2] (7-/5-tap), sgr[3]
    (5x5, 3x3, mix). fn(dst, stride, left, lpf, w, h, params, edges); `params` = address of a
    16-byte aligned LooprestorationParams."""

    def __init__(self, bpc, lib=None):
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 5)()
        (self.lib.b200_loop_restoration_dsp_init_8bpc if bpc == 8 else self.lib.b200_loop_restoration_dsp_init_16bpc)(self._tbl, bpc)
        t = wrap_dsp_table(self._tbl, [("wiener", 2), ("sgr", 3)], {"wiener": (LR_PROTO, True), "sgr": (LR_PROTO, True)},
                           bpc > 8, self.bitdepth_max)
        self.wiener, self.sgr = t["wiener"], t["sgr"]


IPRED_LAYOUT = [("intra_pred", 14), ("cfl_ac", 3), ("cfl_pred", 6), ("pal_pred", 1)]
IPRED_PROTOS = {   # reference src/ipred.h:44-79
    "intra_pred": ([_P, _S, _P, _I, _I, _I, _I, _I], True),
    "cfl_ac": ([_P, _P, _S, _I, _I, _I, _I], False),
    "cfl_pred": ([_P, _S, _P, _I, _I, _P, _I], True),
    "pal_pred": ([_P, _S, _P, _P, _I, _I], False),
}


class IntraPredDSPContext:
    """Dav1dIntraPredDSPContext (reference src/ipred.h:81-90): intra_pred[14], cfl_ac[3], cfl_pred[6] (slots
    DC / LEFT_DC / TOP_DC / DC_128 used), pal_pred."""

    def __init__(self, bpc, lib=None):
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 24)()
        (self.lib.b200_intra_pred_dsp_init_8bpc if bpc == 8 else self.lib.b200_intra_pred_dsp_init_16bpc)(self._tbl)
        for k, v in wrap_dsp_table(self._tbl, IPRED_LAYOUT, IPRED_PROTOS, bpc > 8, self.bitdepth_max).items():
            setattr(self, k, v)


FG_LAYOUT = [("generate_grain_y", 1), ("generate_grain_uv", 3), ("fgy_32x32xn", 1), ("fguv_32x32xn", 3)]
FG_PROTOS = {   # reference src/filmgrain.h:46-73
    "generate_grain_y": ([_P, _P], True),
    "generate_grain_uv": ([_P, _P, _P, C.c_ssize_t], True),
    "fgy_32x32xn": ([_P, _P, _S, _P, C.c_size_t, _P, _P, _I, _I], True),
    "fguv_32x32xn": ([_P, _P, _S, _P, C.c_size_t, _P, _P, _I, _I, _P, _S, _I, _I], True),
}


class FilmGrainDSPContext:
    """Dav1dFilmGrainDSPContext (reference src/filmgrain.h:75-80): generate_grain_y, generate_grain_uv[3]
    (420 / 422 / 444), fgy_32x32xn, fguv_32x32xn[3]. Film grain parameter blocks are passed as the address of
    a _lib.FilmGrainData (ctypes.addressof)."""

    def __init__(self, bpc, lib=None):
        self.bpc, self.bitdepth_max = bpc, (1 << bpc) - 1
        self.lib = lib or get_lib()
        self._tbl = (C.c_void_p * 8)()
        (self.lib.b200_film_grain_dsp_init_8bpc if bpc == 8 else self.lib.b200_film_grain_dsp_init_16bpc)(self._tbl)
        for k, v in wrap_dsp_table(self._tbl, FG_LAYOUT, FG_PROTOS, bpc > 8, self.bitdepth_max).items():
            setattr(self, k, v)
