"""Streams with chosen statistics — TEST INFRASTRUCTURE (lives under tests/, needs oracle/_ref/libdav1d_gen.so; the product
package never imports it).

`obu.py` writes valid headers and fills the tiles with random bytes; an arithmetic decoder fed random bits draws every symbol
from its context's CDF, so such streams have the statistics of AV1's default CDFs: about half of the blocks carry a residual
and every coded transform block is dense. `generate()` makes streams whose block skip rate, intra share and coefficient
sparsity are chosen instead, without an AV1 encoder: the reference decoder itself, built with its symbol decoder replaced
(oracle/gen/gen_msac.c), parses the placeholder stream, CHOOSES every symbol (from the CDF, or by the policy for a few
syntax elements) and range-ENCODES its choices; the tile payloads it leaves behind go into the very same headers.
Self-check: stock dav1d decodes the result to the pictures the generator run reconstructed from its own choices."""
import ctypes as C
import os

import numpy as np

from dav1d_b200 import obu, stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN_SO = os.path.join(ROOT, "oracle", "_ref", "libdav1d_gen.so")
_dll = None


def have_generator():
    return os.path.exists(GEN_SO)


def _gen():
    global _dll
    if _dll is None:
        _dll = C.CDLL(GEN_SO)
        _dll.gen_reset.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int]
        _dll.gen_tile.restype = C.c_uint64
        _dll.gen_tile.argtypes = [C.c_int, C.c_void_p, C.c_uint64]
    return _dll


def generate(build, seed=1, p_skip=-1.0, p_intra=-1.0, p_txskip=-1.0, eob_draws=1, check=True, apply_grain=0, tries=20, layout422=False):
    """build() -> list of temporal units (a call of obu.inter_stream / obu.intra_stream with fixed arguments: it is called twice
    and must make the same header choices both times). Policy (a value < 0 = leave it to the CDF): p_skip = share of skipped
    blocks, p_intra = share of intra blocks in inter frames, p_txskip = share of all-zero transform blocks among the coded ones,
    eob_draws = k: the end-of-block position is the smallest of k draws (sparser coefficients). layout422: the streams are
    4:2:2 — the generator then never chooses the partitions that are illegal there (random payloads hit them all the time).
    Returns (temporal units, n_pictures, info, packed pictures of the generator run)."""
    g = _gen()
    placeholder = build()
    g.gen_set_422(1 if layout422 else 0)
    for attempt in range(tries):
        g.gen_reset(seed + 7919 * attempt, p_skip, p_intra, p_txskip, eob_draws)
        # one thread, no frame delay: tiles are parsed in stream order, frame after frame
        r, info, packed = stream.decode_stream(g, placeholder, n_threads=1, max_frame_delay=1, apply_grain=apply_grain)
        n_tiles = g.gen_finish()
        if r > 0:
            break
        # a chosen symbol made the frame illegal (e.g. an intra block copy vector into the current superblock): other choices
    else:
        raise RuntimeError("the generator found no legal set of choices in %d tries (last error %d)" % (tries, r))
    payloads = []
    for i in range(n_tiles):
        n = g.gen_tile(i, None, 0)
        buf = (C.c_uint8 * n)()
        assert g.gen_tile(i, buf, n) == n
        payloads.append(bytes(buf))
    obu.PAYLOADS = iter(payloads)
    try:
        tus = build()
        leftover = sum(1 for _ in obu.PAYLOADS)
    finally:
        obu.PAYLOADS = None
    assert leftover == 0, "the rebuilt stream has %d tiles fewer than the generator parsed" % leftover
    if check:
        import refs
        r2, info2, packed2 = stream.decode_stream(C.CDLL(refs.REF_SO), tus, n_threads=2, max_frame_delay=2, apply_grain=apply_grain)
        if not (r2 == r and np.array_equal(info, info2) and np.array_equal(packed, packed2)):
            raise RuntimeError("generated stream does not decode to the generator's own pictures (%d vs %d frames)" % (r2, r))
    return tus, r, info, packed
