"""GPU tests of what round 2 finished after its GPU budget was spent (kernel paths verified on the CPU emulator only, behind
the real dav1d front end where streams are involved). They live in the file pytest runs last, so that with `-x` everything
that has been GPU-verified before is still run and reported first."""
import numpy as np
import pytest

import refs
from dav1d_b200 import obu, stream, synth
from test_intra import IBC_CASES, oracle_intra, planes_equal, run_lib
from test_stream import _check


@pytest.fixture(scope="module")
def gpu_decoder():
    d = stream.HookedDecoder()
    yield d
    d.release()


@pytest.mark.gpu
@pytest.mark.parametrize("bpc,W,H,ssh,ssv", IBC_CASES[:2] + [(8, 1288, 720, 1, 1)])
def test_gpu_intra_block_copy_records(bpc, W, H, ssh, ssv):
    S = synth.make_intra_frame(np.random.default_rng(760 + bpc + W), bpc, W, H, ssh, ssv, p_ibc=0.3)
    exp = oracle_intra(S)
    got = run_lib(None, None, S)
    ok, where = planes_equal(S, exp, got)
    assert ok, where



@pytest.mark.gpu
def test_scaled_references_and_super_resolution_gpu(gpu_decoder):
    """on the device: frames coded at changing sizes (scaled predictions against per-reference geometry) and super-resolution
    (resize stage before loop restoration; later frames predict from the upscaled pictures), key and inter frames"""
    n_scaled = 0
    for seed, (w, h, sizes, kw) in enumerate([(256, 192, [(192, 144), (256, 192), (160, 96)], dict(bpc=8)),
                                              (320, 192, [(256, 160), (320, 192), (200, 120), (320, 176)], dict(bpc=10, motion_modes=1, film_grain=1))]):
        tus = obu.inter_stream(700 + seed, w, h, n_frames=6, sizes=sizes, **kw)
        _check(gpu_decoder, tus, 6, apply_grain=1)
        n_scaled += gpu_decoder.last_stats["scaled"]
    for seed, (w, h, kw) in enumerate([(328, 200, dict(bpc=10, log2_cols=1)), (256, 192, dict(bpc=8, layout="400")), (320, 192, dict(bpc=8, film_grain=1))]):
        _check(gpu_decoder, obu.intra_stream(900 + seed, w, h, n_frames=2, super_res=1, **kw), 2, apply_grain=1)
    for seed, (w, h, kw) in enumerate([(256, 192, dict(bpc=8)), (328, 200, dict(bpc=10, log2_cols=1, motion_modes=1))]):
        tus = obu.inter_stream(950 + seed, w, h, n_frames=6, super_res=1, **kw)
        _check(gpu_decoder, tus, 6, apply_grain=1)
        n_scaled += gpu_decoder.last_stats["scaled"]
    assert n_scaled > 100




@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gen_422_10bit_all_tools", "gen_sparse_8bit", "inter_444_12bit"])
def test_golden_streams_round2_gpu_md5(gpu_decoder, name):
    """generator-made golden streams (4:2:2 at a real frame size with every inter tool and film grain; encoder-like sparse
    statistics) and the 4:4:4 12-bit one: the committed md5 of stock dav1d's output, no reference / oracle needed on the box"""
    from test_stream import _golden, _md5
    tus, want = _golden()[name]
    assert _md5(gpu_decoder.decode(tus, apply_grain=1)) == (want["md5"], want["frames"])
    gpu_decoder.stats(reset=True)
