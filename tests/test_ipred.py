"""Parity tests for intra prediction (Dav1dIntraPredDSPContext), following tests/checkasm/ipred.c:
all 14 modes, w in {4..64}, h in [w/4, 4w], Z angles from z_angles[27] with random 0x600 flag bits, Z2
max_width/height edge cases (:69-78), filter index (:114-115); cfl_ac with all pad combinations (:157-205),
cfl_pred (:207-258), pal_pred (:260-296)."""
import ctypes as C
import numpy as np
import pytest

import refs

Z_ANGLES = [3, 6, 9, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45, 48, 51, 54, 58, 61, 64, 67, 70, 73, 76, 81, 84, 87]
Z1, Z2, Z3, FILTER = 6, 7, 8, 13
PAD = 8


def ref_ipred(bpc):
    from dav1d_b200 import dsp
    t = (C.c_void_p * 24)()
    (refs.ref().dav1d_intra_pred_dsp_init_8bpc if bpc == 8 else refs.ref().dav1d_intra_pred_dsp_init_16bpc)(t)

    class Ctx:
        pass
    c = Ctx(); c._t = t
    for k, v in dsp.wrap_dsp_table(t, dsp.IPRED_LAYOUT, dsp.IPRED_PROTOS, bpc > 8, (1 << bpc) - 1).items():
        setattr(c, k, v)
    return c


def oracle_ipred(bpc):
    o = refs.oracle(); bd = (1 << bpc) - 1
    P, S = C.c_void_p, C.c_ssize_t

    def a(x):
        return x.ctypes.data if isinstance(x, np.ndarray) else x

    class Ctx:
        pass
    c = Ctx()
    c.intra_pred = [(lambda d, st, tl, w, h, ang, mw, mh, m=m: o.oracle_ipred(m, P(a(d)), S(st), P(a(tl)), w, h, ang, mw, mh, bd))
                    for m in range(14)]
    lay = [(1, 1), (1, 0), (0, 0)]
    c.cfl_ac = [(lambda ac, y, st, wp, hp, cw, ch, l=l: o.oracle_cfl_ac(P(a(ac)), P(a(y)), S(st), wp, hp, cw, ch, l[0], l[1], bd)) for l in lay]
    c.cfl_pred = [((lambda d, st, tl, w, h, ac, al, m=m: o.oracle_cfl_pred(m, P(a(d)), S(st), P(a(tl)), w, h, P(a(ac)), al, bd))
                   if m in (0, 3, 4, 5) else None) for m in range(6)]
    c.pal_pred = lambda d, st, pal, idx, w, h: o.oracle_pal_pred(P(a(d)), S(st), P(a(pal)), P(a(idx)), w, h, bd)
    return c


def gen_z2_max(rng, sz):
    n = int(rng.integers(0, 1 << 20))
    if n & (1 << 17):
        return (n & (sz - 1)) + 1
    if n & (1 << 16):
        return 65536
    return (n & 65535) + 1


def run_ipred_checks(new, chk, bpc, seed, light=False):
    rng = np.random.default_rng(seed)
    bd = (1 << bpc) - 1
    dt = refs.pixel_dtype(bpc)
    isz = np.dtype(dt).itemsize
    n = 0
    for mode in range(14):
        wmax = 32 if mode == FILTER else 64
        w = 4
        while w <= wmax:
            h = max(w // 4, 4)
            while h <= min(w * 4, wmax):
                iters = (2 if light else 5) if Z1 <= mode <= Z3 else 1
                for _ in range(iters):
                    ang = mw = mh = 0
                    if Z1 <= mode <= Z3:
                        ang = (90 * (mode - Z1) + Z_ANGLES[int(rng.integers(0, 27))]) | (int(rng.integers(0, 4)) << 9)
                        if mode == Z2:
                            mw, mh = gen_z2_max(rng, w), gen_z2_max(rng, h)
                    elif mode == FILTER:
                        ang = int(rng.integers(0, 5)) | (int(rng.integers(0, 8)) << 9)
                    edge = np.zeros(257 + 32, dt)
                    tlo = 128 + 16
                    edge[tlo - 2 * h:tlo + 2 * w + 1] = rng.integers(0, bd + 1, 2 * h + 2 * w + 1)
                    c1 = np.zeros((h + 2 * PAD, w + 2 * PAD), dt); c2 = c1.copy()
                    tl = edge[tlo:]
                    chk.intra_pred[mode](c1[PAD:, PAD:], c1.strides[0], tl, w, h, ang, mw, mh)
                    new.intra_pred[mode](c2[PAD:, PAD:], c2.strides[0], tl, w, h, ang, mw, mh)
                    assert np.array_equal(c1, c2), ("intra_pred", bpc, mode, w, h, ang, mw, mh)
                    n += 1
                h <<= 1
            w <<= 1
    # cfl_ac
    for li, (ssh, ssv) in enumerate([(1, 1), (1, 0), (0, 0)]):
        hs, vs = 2 >> ssh, 2 >> ssv
        w = 4
        while w <= (32 >> ssh):
            h = max(w // 4, 4)
            while h <= min(w * 4, 32 >> ssv):
                wp = max((w >> 2) - hs, 0)
                while wp >= 0:
                    hp = max((h >> 2) - vs, 0)
                    while hp >= 0:
                        luma = rng.integers(0, bd + 1, (32, 32)).astype(dt)
                        a1 = np.zeros(32 * 32, np.int16); a2 = a1.copy()
                        chk.cfl_ac[li](a1, luma, 32 * isz, wp, hp, w, h)
                        new.cfl_ac[li](a2, luma, 32 * isz, wp, hp, w, h)
                        assert np.array_equal(a1, a2), ("cfl_ac", bpc, li, w, h, wp, hp)
                        n += 1
                        hp -= vs
                    wp -= hs
                h <<= 1
            w <<= 1
    # cfl_pred
    for mode in (0, 3, 4, 5):
        w = 4
        while w <= 32:
            h = max(w // 4, 4)
            while h <= min(w * 4, 32):
                alpha = (int(rng.integers(0, 16)) + 1) * (1 - (int(rng.integers(0, 4)) & 2))
                edge = np.zeros(257 + 32, dt); tlo = 128 + 16
                edge[tlo - 2 * h:tlo + 2 * w + 1] = rng.integers(0, bd + 1, 2 * h + 2 * w + 1)
                acv = rng.integers(0, (bd << 3) + 1, w * h).astype(np.int64)
                avg = (int(acv.sum()) + (w * h >> 1)) // (w * h)
                ac = (acv - avg).astype(np.int16)
                c1 = np.zeros((h + 2 * PAD, w + 2 * PAD), dt); c2 = c1.copy()
                chk.cfl_pred[mode](c1[PAD:, PAD:], c1.strides[0], edge[tlo:], w, h, ac, alpha)
                new.cfl_pred[mode](c2[PAD:, PAD:], c2.strides[0], edge[tlo:], w, h, ac, alpha)
                assert np.array_equal(c1, c2), ("cfl_pred", bpc, mode, w, h, alpha)
                n += 1
                h <<= 1
            w <<= 1
    # pal_pred
    w = 4
    while w <= 64:
        h = max(w // 4, 4)
        while h <= min(w * 4, 64):
            pal = rng.integers(0, bd + 1, 8).astype(dt)
            idx = (rng.integers(0, 256, w * h // 2) & 0x77).astype(np.uint8)
            c1 = np.zeros((h + 2 * PAD, w + 2 * PAD), dt); c2 = c1.copy()
            chk.pal_pred(c1[PAD:, PAD:], c1.strides[0], pal, idx, w, h)
            new.pal_pred(c2[PAD:, PAD:], c2.strides[0], pal, idx, w, h)
            assert np.array_equal(c1, c2), ("pal_pred", bpc, w, h)
            n += 1
            h <<= 1
        w <<= 1
    return n


@pytest.mark.parametrize("bpc", [8, 10, 12])
def test_oracle_ipred_vs_reference(bpc):
    if not refs.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    n = 0
    for rep in range(3):
        n += run_ipred_checks(oracle_ipred(bpc), ref_ipred(bpc), bpc, seed=700 + bpc + rep)
    assert n > 1500


@pytest.mark.emu
@pytest.mark.parametrize("bpc", [8, 10])
def test_emu_ipred(bpc):
    from dav1d_b200.dsp import IntraPredDSPContext
    run_ipred_checks(IntraPredDSPContext(bpc, lib=refs.emu_lib()), oracle_ipred(bpc), bpc, seed=710 + bpc, light=True)


@pytest.mark.gpu
@pytest.mark.parametrize("bpc", [8, 10, 12])
def test_gpu_ipred(bpc):
    from dav1d_b200.dsp import IntraPredDSPContext
    chk = ref_ipred(bpc) if refs.have_ref() else oracle_ipred(bpc)
    new = IntraPredDSPContext(bpc)
    assert run_ipred_checks(new, chk, bpc, seed=720 + bpc) > 500
    run_ipred_checks(new, oracle_ipred(bpc), bpc, seed=730 + bpc, light=True)
