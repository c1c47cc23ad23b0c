"""The stream generator (tests/streamgen.py, oracle/gen/gen_msac.c): the reference decoder with its symbol decoder replaced
by one that chooses and range-encodes the symbols. Its streams must decode, with stock dav1d, to the pictures the generator
itself reconstructed (the self-check inside generate()), follow the statistics asked for, and — like every stream — decode
identically through the hooked decoder. 4:2:2 at real frame sizes exists only this way."""
import importlib.util
import os

import numpy as np
import pytest

import refs
import streamgen
from dav1d_b200 import obu, stream
from test_stream import _check

pytestmark = pytest.mark.skipif(not streamgen.have_generator() or not refs.have_ref(), reason="oracle/_ref/libdav1d_gen.so not built")


@pytest.fixture(scope="module")
def emu_decoder():
    refs.emu_lib()
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(refs.ROOT, "tests", "emu", "build_emu.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    d = stream.HookedDecoder(backend=m.build(), serialize=True)
    yield d
    d.release()


def _build(kind, seed, w, h, **kw):
    return (lambda: obu.inter_stream(seed, w, h, **kw)) if kind == "inter" else (lambda: obu.intra_stream(seed, w, h, **kw))


@pytest.mark.parametrize("seed", range(8))
def test_generated_streams_decode_to_the_generators_pictures(seed):
    """natural sampling (every symbol drawn from its CDF, what a random payload does): the range encoder is the exact inverse
    of the decoder — stock dav1d reproduces the generator's pictures — across bit depths, layouts, tiles, tools"""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(16, 70)) * 8 + int(rng.choice([0, 2, 6])), int(rng.integers(12, 44)) * 8 + int(rng.choice([0, 4]))
    kw = dict(bpc=int(rng.choice([8, 10, 12])), log2_cols=int(rng.integers(0, 3)), log2_rows=int(rng.integers(0, 2)),
              film_grain=int(rng.integers(0, 2)), layout=str(rng.choice(["420", "444", "400"])), sb128=int(rng.integers(0, 2)))
    if seed % 2:
        kw.update(n_frames=5, motion_modes=2, hidden_every=int(rng.choice([0, 2])), segmentation=int(rng.integers(0, 2)))
    else:
        kw.update(n_frames=2, screen_content=int(rng.integers(0, 2)))
    tus, n, _, _ = streamgen.generate(_build("inter" if seed % 2 else "intra", 100 + seed, w, h, **kw), seed=seed, apply_grain=1)
    assert n >= kw["n_frames"] and (kw.get("hidden_every") or n == kw["n_frames"])     # hidden frames come out a second time
    assert len(b"".join(tus)) < len(b"".join(_build("inter" if seed % 2 else "intra", 100 + seed, w, h, **kw)()))


def test_policy_changes_the_statistics(emu_decoder):
    """mostly skipped blocks, sparse coefficients, few intra blocks: counted through the hooked decoder's record statistics"""
    build = _build("inter", 7, 640, 360, n_frames=5, motion_modes=2, log2_cols=1)
    per_frame = {}
    for name, pol in (("natural", {}), ("sparse", dict(p_skip=0.85, p_txskip=0.7, eob_draws=8, p_intra=0.05)),
                      ("intra-heavy", dict(p_intra=0.6)), ("intra-light", dict(p_intra=0.02))):
        tus, n, _, _ = streamgen.generate(build, seed=3, **pol)
        _check(emu_decoder, tus, n)
        st = emu_decoder.last_stats
        per_frame[name] = (st["coefs"] / st["frames"], st["intra_tx"] / st["frames"], len(b"".join(tus)))
    assert per_frame["sparse"][0] < 0.25 * per_frame["natural"][0], per_frame          # staged coefficients
    assert per_frame["intra-light"][1] < 0.6 * per_frame["intra-heavy"][1], per_frame  # intra transform blocks (key frame included)
    assert per_frame["sparse"][2] < 0.4 * per_frame["natural"][2], per_frame           # bytes


@pytest.mark.parametrize("kind,w,h,kw", [("intra", 640, 360, dict(bpc=8, n_frames=2)),
                                         ("inter", 704, 480, dict(bpc=10, n_frames=5, motion_modes=2, film_grain=1, log2_cols=1)),
                                         ("inter", 416, 240, dict(bpc=12, n_frames=4, motion_modes=1, sizes=[(320, 192), (416, 240)]))])
def test_422_at_real_frame_sizes(emu_decoder, kind, w, h, kw):
    """4:2:2 streams of real sizes (random payloads hit illegal partitions at once: the generator never chooses them) through
    the hooked decoder, byte-identical to stock dav1d"""
    tus, n, _, _ = streamgen.generate(_build(kind, 9, w, h, layout="422", **kw), seed=2, layout422=True, p_skip=0.5, eob_draws=3, apply_grain=1)
    _check(emu_decoder, tus, n, apply_grain=1)
