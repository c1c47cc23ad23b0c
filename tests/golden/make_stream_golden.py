"""Regenerates tests/golden/stream_golden.json + the .obu files next to it: small synthetic AV1 streams (dav1d_b200/obu.py)
and the md5 of the pictures the UNMODIFIED reference decoder (oracle/_ref/libdav1d_ref.so, built from /root/reference by
oracle/Makefile) produces for them, film grain applied. Run where the reference is built:

    python tests/golden/make_stream_golden.py
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dav1d_b200 import cli, obu, stream  # noqa: E402
import streamgen                          # noqa: E402


def _gen(build, **policy):
    """a stream whose symbols the reference decoder chose and range-encoded (tests/streamgen.py)"""
    return lambda: streamgen.generate(build, apply_grain=1, **policy)[0]


CASES = {
    "key_8bit_tiles": lambda: obu.intra_stream(1, 136, 96, n_frames=2, log2_cols=1, payload_bytes_per_sb64=700),
    "inter_10bit_all_tools": lambda: obu.inter_stream(2, 136, 96, n_frames=4, bpc=10, motion_modes=2, film_grain=1, screen_content=1,
                                                      global_motion=1, segmentation=1, hidden_every=2, payload_bytes_per_sb64=700),
    "inter_444_12bit": lambda: obu.inter_stream(3, 72, 72, n_frames=3, bpc=12, layout="444", motion_modes=1, payload_bytes_per_sb64=1500),
    # round 2, made by the stream generator: 4:2:2 at a real frame size (random payloads are illegal there) and a stream with
    # encoder-like statistics (85 % skipped blocks, sparse coefficients, 5 % intra blocks)
    "gen_422_10bit_all_tools": _gen(lambda: obu.inter_stream(9, 416, 240, n_frames=5, bpc=10, layout="422", motion_modes=2, film_grain=1, log2_cols=1),
                                    seed=2, layout422=True, p_skip=0.5, eob_draws=3),
    "gen_sparse_8bit": _gen(lambda: obu.inter_stream(7, 640, 360, n_frames=6, motion_modes=2, log2_cols=1, segmentation=1),
                            seed=3, p_skip=0.85, p_txskip=0.7, eob_draws=8, p_intra=0.05),
}


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdav1d_ref.so"))
    out = {}
    for name, gen in CASES.items():
        tus = gen()
        with open(os.path.join(HERE, "stream_%s.obu" % name), "wb") as fh:
            fh.write(b"".join(tus))
        n, info, packed = stream.decode_stream(ref, tus, apply_grain=1)
        assert n > 0, (name, n)
        digest, frames = cli.md5_of(cli.frames_of(info, packed))
        out[name] = {"md5": digest, "frames": frames, "temporal_units": len(tus), "bytes": sum(map(len, tus))}
        print(name, out[name])
    json.dump(out, open(os.path.join(HERE, "stream_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
