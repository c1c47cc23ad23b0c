"""Parity tests for the inverse transforms (Dav1dInvTxfmDSPContext).

Modelled on the reference's own test body, tests/checkasm/itx.c:245-310: for every defined
itxfm_add[tx][txtp] slot and every `subsh` sub-block class, coefficients come from the float
forward-transform generator, dst is random, and BOTH the destination rectangle (with padding
guards) and the coefficient buffer after the call (the zeroing contract) must be identical.

  not gpu : the plain-C oracle against the UNMODIFIED reference C path and against the committed
            golden fixtures (this is what pins the oracle); the CUDA sources on the host emulator
  gpu     : the CUDA kernels through the C ABI (Level-1 tables and Level-2 batches) against the
            reference C path (when oracle/_ref is present) and the oracle. Bit-exact.
"""
import os
import numpy as np
import pytest

import refs
from dav1d_b200 import levels as L

PAD = 16  # guard columns/rows around the destination rectangle (checkasm's padded-rect check)


def slots(bpc_list=(8, 10, 12), every=1):
    k = 0
    for bpc in bpc_list:
        for tx in range(L.N_RECT_TX_SIZES):
            w, h = L.TX_W[tx], L.TX_H[tx]
            smax = refs.SUBSH_ITERS[int(np.log2(max(w, h))) - 2]
            for tp in range(L.N_TX_TYPES_PLUS_LL):
                if not L.itx_defined(tx, tp):
                    continue
                for subsh in range(1 if tp else 0, smax):
                    k += 1
                    if k % every == 0:
                        yield bpc, tx, tp, subsh


def run_checkasm_itx(new_tbls, ref_tbls, cases, seed, neg_stride_every=7):
    """new_tbls / ref_tbls: {bpc: itxfm_add[tx][txtp]} of callables (dst, stride_bytes, coeff, eob)."""
    rng = np.random.default_rng(seed)
    n = 0
    for bpc, tx, tp, subsh in cases:
        bdmax = (1 << bpc) - 1
        w, h = L.TX_W[tx], L.TX_H[tx]
        coef, eob = refs.gen_itx_coefs(rng, tx, tp, subsh, bdmax)
        # garbage after the coded block, as the reference test leaves it (itx.c:180-181)
        cbuf = rng.integers(-32768, 32767, 32 * 32).astype(refs.coef_dtype(bpc))
        cbuf[:len(coef)] = coef.astype(refs.coef_dtype(bpc))
        canvas = rng.integers(0, bdmax + 1, (h + 2 * PAD, w + 2 * PAD)).astype(refs.pixel_dtype(bpc))
        c_ref, c_new = canvas.copy(), canvas.copy()
        k_ref, k_new = cbuf.copy(), cbuf.copy()
        n += 1
        if n % neg_stride_every == 0:   # bottom-up picture: negative stride (dav1d --negstride)
            d_ref, d_new = c_ref[PAD + h - 1:, PAD:], c_new[PAD + h - 1:, PAD:]
            stride = -canvas.strides[0]
            # with a negative stride row y of the block is canvas row PAD+h-1-y
        else:
            d_ref, d_new = c_ref[PAD:, PAD:], c_new[PAD:, PAD:]
            stride = canvas.strides[0]
        ref_tbls[bpc][tx][tp](d_ref, stride, k_ref, eob)
        new_tbls[bpc][tx][tp](d_new, stride, k_new, eob)
        what = "%dbpc %s %s subsh=%d eob=%d" % (bpc, L.TX_NAMES[tx], L.TXTP_NAMES[tp], subsh, eob)
        assert np.array_equal(c_ref, c_new), "dst mismatch: " + what
        assert np.array_equal(k_ref, k_new), "coef (zeroing contract) mismatch: " + what
    return n


# ------------------------------------------------------------------ oracle pinning (CPU)
@pytest.mark.parametrize("bpc", [8, 10, 12])
def test_oracle_vs_reference_checkasm(bpc):
    if not refs.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    n = run_checkasm_itx({bpc: refs.oracle_itxfm_add(bpc)}, {bpc: refs.ref_itx_table(bpc)},
                         list(slots((bpc,))) * 2, seed=100 + bpc)
    assert n == 592


@pytest.mark.parametrize("bpc", [8, 12])
def test_oracle_vs_reference_garbage(bpc):
    """Out-of-contract inputs: full-range coefficients everywhere and arbitrary eob — the
    restatement must still agree (wrap-around arithmetic, eob-derived zero rows)."""
    if not refs.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    rng = np.random.default_rng(7 + bpc)
    bdmax = (1 << bpc) - 1
    rt, ot = refs.ref_itx_table(bpc), refs.oracle_itxfm_add(bpc)
    amp = 32767 if bpc == 8 else (1 << 19)
    for tx in range(19):
        w, h = L.TX_W[tx], L.TX_H[tx]
        sw, sh = L.tx_coef_dims(tx)
        for tp in range(17):
            if not L.itx_defined(tx, tp):
                continue
            for it in range(6):
                cf = rng.integers(-amp, amp + 1, sw * sh).astype(refs.coef_dtype(bpc))
                eob = int(rng.integers(0, sw * sh))
                dst = rng.integers(0, bdmax + 1, (h, w)).astype(refs.pixel_dtype(bpc))
                d1, d2, c1, c2 = dst.copy(), dst.copy(), cf.copy(), cf.copy()
                rt[tx][tp](d1, d1.strides[0], c1, eob)
                ot[tx][tp](d2, d2.strides[0], c2, eob)
                assert np.array_equal(d1, d2) and np.array_equal(c1, c2), (bpc, tx, tp, eob)


def golden_cases():
    g = np.load(os.path.join(refs.ROOT, "tests", "golden", "itx_golden.npz"))
    for bpc in (8, 10, 12):
        for tx in range(19):
            for tp in range(17):
                k = "b%d_t%d_p%d_" % (bpc, tx, tp)
                if k + "coef" in g:
                    yield bpc, tx, tp, g[k + "coef"], int(g[k + "eob"]), g[k + "dst"], g[k + "out"], g[k + "coef_out"]


def check_golden(tbls):
    n = 0
    for bpc, tx, tp, coef, eob, dst, out, coef_out in golden_cases():
        d, c = dst.copy(), coef.copy()
        tbls[bpc][tx][tp](d, d.strides[0], c, eob)
        assert np.array_equal(d, out) and np.array_equal(c, coef_out), (bpc, tx, tp)
        n += 1
    assert n == 156 * 3
    return n


def test_oracle_golden_fixtures():
    """Committed vectors produced by the reference (tests/golden/make_golden.py)."""
    check_golden({bpc: refs.oracle_itxfm_add(bpc) for bpc in (8, 10, 12)})


# ------------------------------------------------------------------ host emulator (debug harness)
@pytest.mark.emu
def test_emu_itx_golden():
    from dav1d_b200.dsp import InvTxfmDSPContext
    lib = refs.emu_lib()
    check_golden({bpc: InvTxfmDSPContext(bpc, lib=lib).itxfm_add for bpc in (8, 10, 12)})


@pytest.mark.emu
def test_emu_itx_checkasm_subset():
    from dav1d_b200.dsp import InvTxfmDSPContext
    lib = refs.emu_lib()
    new = {bpc: InvTxfmDSPContext(bpc, lib=lib).itxfm_add for bpc in (8, 10, 12)}
    orc = {bpc: refs.oracle_itxfm_add(bpc) for bpc in (8, 10, 12)}
    run_checkasm_itx(new, orc, list(slots(every=5)), seed=5)


# ------------------------------------------------------------------ GPU parity (through the C ABI)
def _checkers():
    """reference C path when shipped, plus the oracle"""
    out = [("oracle", {bpc: refs.oracle_itxfm_add(bpc) for bpc in (8, 10, 12)})]
    if refs.have_ref():
        out.append(("reference", {bpc: refs.ref_itx_table(bpc) for bpc in (8, 10, 12)}))
    return out


@pytest.mark.gpu
def test_gpu_itx_level1_all_slots():
    from dav1d_b200.dsp import InvTxfmDSPContext
    new = {bpc: InvTxfmDSPContext(bpc).itxfm_add for bpc in (8, 10, 12)}
    for bpc in (8, 10, 12):          # the table has exactly dav1d's 156 non-NULL slots
        assert sum(f is not None for row in new[bpc] for f in row) == 156
    check_golden(new)
    for name, chk in _checkers():
        run_checkasm_itx(new, chk, list(slots()), seed=11)


def make_batch(rng, bpc, tx, n, types=None, plane_w=None):
    """n non-overlapping blocks of size tx tiled over a random picture; checkasm-style coefs."""
    bdmax = (1 << bpc) - 1
    w, h = L.TX_W[tx], L.TX_H[tx]
    sw, sh = L.tx_coef_dims(tx)
    types = types or [tp for tp in range(17) if L.itx_defined(tx, tp)]
    per_row = max(1, int(np.ceil(np.sqrt(n * h / w))))
    rows = (n + per_row - 1) // per_row
    stride = per_row * w + 24
    pic = rng.integers(0, bdmax + 1, (rows * h + 3, stride)).astype(refs.pixel_dtype(bpc))
    blocks = np.zeros(n, refs.ITX_BLOCK_DTYPE)
    coefs = np.zeros(n * sw * sh, refs.coef_dtype(bpc))
    smax = refs.SUBSH_ITERS[int(np.log2(max(w, h))) - 2]
    order = rng.permutation(n)
    for i in range(n):
        tp = types[int(rng.integers(0, len(types)))]
        subsh = int(rng.integers(1 if tp else 0, smax))
        c, eob = refs.gen_itx_coefs(rng, tx, tp, subsh, bdmax)
        slot = int(order[i])
        coefs[slot * sw * sh:(slot + 1) * sw * sh] = c
        by, bx = divmod(i, per_row)
        blocks[i] = (by * h * stride + bx * w + 5, slot * sw * sh, eob, tp, 0)
    return blocks, coefs, pic, stride


@pytest.mark.gpu
@pytest.mark.parametrize("bpc", [8, 10, 12])
def test_gpu_itx_batch_all_sizes(bpc):
    import torch
    from dav1d_b200 import batch
    rng = np.random.default_rng(21 + bpc)
    bdmax = (1 << bpc) - 1
    for tx in range(19):
        n = 257 if max(L.TX_W[tx], L.TX_H[tx]) >= 32 else 1031
        blocks, coefs, pic, stride = make_batch(rng, bpc, tx, n)
        exp_pic, exp_coef = pic.copy(), coefs.copy()
        st = (refs.C.c_int32 * 3)(stride, stride, stride)
        assert refs.oracle().oracle_itx_add_batch(bdmax, tx, blocks.ctypes.data, n, exp_coef.ctypes.data,
                                                  exp_pic.ctypes.data, st, 1) == 0
        for zero in (1, 0):
            d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
            d_coef = torch.from_numpy(coefs.copy()).cuda()
            d_pic = torch.from_numpy(pic.copy().view(np.int16 if bpc > 8 else np.uint8)).cuda()
            batch.itx_add_batch(bdmax, tx, d_blocks, d_coef, d_pic, [stride] * 3, zero_coefs=bool(zero))
            torch.cuda.synchronize()
            got = d_pic.cpu().numpy().view(pic.dtype)
            assert np.array_equal(got, exp_pic), "pic mismatch tx=%s bpc=%d" % (L.TX_NAMES[tx], bpc)
            gc = d_coef.cpu().numpy()
            assert np.array_equal(gc, exp_coef if zero else coefs), "coef mismatch tx=%s" % L.TX_NAMES[tx]
        # host-buffer entry point (the e2e path of bench.py)
        p2, c2 = pic.copy(), coefs.copy()
        batch.itx_add_batch_host(bdmax, tx, blocks, c2, p2, [stride] * 3, zero_coefs=True)
        assert np.array_equal(p2, exp_pic) and np.array_equal(c2, exp_coef)
