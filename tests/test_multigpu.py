"""The N>1 path on CPU: two gloo ranks shard a 4-frame GOP (frame n -> rank n mod 2) through the emulated
library, broadcasting every restored picture as the next frames' reference; both ranks must end with the same
pictures as a single-rank decode, and frame 0 must equal the oracle's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import refs

W, H, N = 136, 72, 4


def _frames():
    from dav1d_b200 import synth
    return [synth.make_inter_frame(np.random.default_rng(900 + k), 8, W, H) for k in range(N)]


def _decode(rank, world):
    from dav1d_b200 import frame, shard
    lib = refs.emu_lib()

    def make(S):
        return frame.FrameBuffers(S, lib=lib, alloc=frame.NumpyAlloc())

    def as_tensor(fb, name):
        return torch.from_numpy(fb.keep[name][0])
    return shard.decode_gop(_frames(), make, dist, rank, world, as_tensor)


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pics = _decode(rank, world)
        np.save(os.path.join(outdir, "r%d.npy" % rank), np.stack(pics))
    finally:
        dist.destroy_process_group()


@pytest.mark.emu
def test_two_ranks_shard_a_gop(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "r0.npy"); b = np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b), "ranks disagree"
    single = np.stack(_decode(0, 1))
    assert np.array_equal(a, single), "sharded decode differs from single-rank decode"
    # frame 0 against the oracle
    import test_frame as TF
    import test_looprestoration as TLR
    S0 = _frames()[0]
    exp = TF.oracle_frame(S0)
    assert TLR.picture_equal(S0, a[0], exp["lr"])
    # later frames really depend on the exchanged pictures
    from dav1d_b200 import frame
    lone = frame.FrameBuffers(_frames()[1], lib=refs.emu_lib(), alloc=frame.NumpyAlloc())
    lone.run()
    assert not np.array_equal(lone.output(), a[1])


def test_frame_owner_round_robin():
    from dav1d_b200 import shard
    assert [shard.frame_owner(n, 4) for n in range(6)] == [0, 1, 2, 3, 0, 1]
