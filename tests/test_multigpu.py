"""The N>1 path: a dependent group of pictures sharded over ranks (frame n -> rank n mod world, predicting from the restored
pictures of frames n-1 and n-2 that OTHER ranks produce), band by band, gated like dav1d's check_tile (dav1d_b200/shard.py).

CPU: gloo ranks drive the emulated library, reference rows travel as messages (DistExchange).
GPU: one process per rank and GPU, reference rows travel as puts into peer memory mapped with CUDA IPC, gated by
     stream-ordered flags (PeerExchange); needs as many GPUs as ranks (tools/gpu_r2.sh runs it under gpurun --gpus N).
Every rank's pictures must equal a single-rank decode of the same GOP and the oracle's chained decode."""
import os
import sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # inherited by the spawned ranks (see dav1d_b200/__init__.py)

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import refs

W, H, N = 136, 200, 6


def _frames(bpc=8, w=W, h=H, n=N, seed=900):
    from dav1d_b200 import synth
    return [synth.make_inter_frame(np.random.default_rng(seed + k), bpc, w, h, film_grain=False) for k in range(n)]


def oracle_gop(frames, n_refs=2):
    """the chained decode on the CPU: frame k's references are the restored pictures of frames k-1, k-2"""
    import test_frame as TF
    out = []
    for k, S in enumerate(frames):
        S2 = dict(S)
        S2["refs"] = [out[k - 1 - d] if k - 1 - d >= 0 else S["refs"][d] for d in range(n_refs)]
        out.append(TF.oracle_frame(S2)["lr"])
    return out


def _decode_emu(rank, world, frames, band_rows=64):
    from dav1d_b200 import frame, shard
    lib = refs.emu_lib()

    def make(S, rows):
        return frame.FrameBuffers(S, lib=lib, alloc=frame.NumpyAlloc(), band_rows=rows)

    def new_buffer(nbytes):
        a = np.zeros(nbytes, np.uint8)
        return a, a.ctypes.data
    return shard.decode_gop(frames, make, dist if world > 1 else None, rank, world, lib, exchange="dist", band_rows=band_rows,
                            as_tensor=torch.from_numpy, new_buffer=new_buffer)


def _worker_emu(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pics = _decode_emu(rank, world, _frames())
        np.savez(os.path.join(outdir, "r%d.npz" % rank), **{str(k): v for k, v in pics.items()})
    finally:
        dist.destroy_process_group()


def _collect(outdir, world, n):
    got = {}
    for r in range(world):
        z = np.load(os.path.join(outdir, "r%d.npz" % r))
        for k in z.files:
            assert int(k) % world == r and int(k) not in got
            got[int(k)] = z[k]
    assert sorted(got) == list(range(n))
    return [got[k] for k in range(n)]


@pytest.mark.emu
@pytest.mark.parametrize("world", [2, 3, 4])
def test_ranks_shard_a_dependent_gop(tmp_path, world):
    port = 29500 + (os.getpid() + world * 7) % 2000
    mp.spawn(_worker_emu, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    frames = _frames()
    got = _collect(str(tmp_path), world, N)
    single = _decode_emu(0, 1, frames)
    import test_looprestoration as TLR
    exp = oracle_gop(frames)
    for k in range(N):
        assert np.array_equal(got[k], single[k]), "frame %d: sharded decode differs from the single-rank decode" % k
        assert TLR.picture_equal(frames[k], got[k], exp[k]), "frame %d differs from the oracle's chained decode" % k
    # later frames really depend on the exchanged pictures
    from dav1d_b200 import frame
    lone = frame.FrameBuffers(frames[1], lib=refs.emu_lib(), alloc=frame.NumpyAlloc())
    lone.run()
    assert not np.array_equal(lone.output(), got[1])


def _mixed_frames():
    from dav1d_b200 import synth
    return [synth.make_inter_frame(np.random.default_rng(40 + k), 8, 200, 136, p_intra=0.15, p_obmc=0.2, p_warp=0.1, p_ii=0.1) for k in range(6)]


def _worker_emu_mixed(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pics = _decode_emu(rank, world, _mixed_frames(), band_rows=192)          # 136 rows: the band is the whole frame
        np.savez(os.path.join(outdir, "r%d.npz" % rank), **{str(k): v for k, v in pics.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.emu
def test_ranks_shard_mixed_frames_as_one_band(tmp_path):
    """frames with intra-machine records (intra blocks, inter-intra) and OBMC / warps over two ranks: such frames are not cut
    into bands (the band is the frame), the exchange and the dependency rule are the same"""
    port = 29500 + (os.getpid() + 31) % 2000
    mp.spawn(_worker_emu_mixed, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    exp = oracle_gop(_mixed_frames())
    got = _collect(str(tmp_path), 2, 6)
    for k, (a, b) in enumerate(zip(exp, got)):
        assert np.array_equal(a, b), "frame %d" % k


@pytest.mark.emu
def test_single_rank_pipeline_two_band_sizes():
    """world = 1: the band pipeline is just a chained decode; 64- and 128-row bands agree with the oracle"""
    import test_looprestoration as TLR
    frames = _frames(n=4)
    exp = oracle_gop(frames)
    for rows in (64, 128):
        got = _decode_emu(0, 1, frames, band_rows=rows)
        for k in range(4):
            assert TLR.picture_equal(frames[k], got[k], exp[k]), (rows, k)


def test_frame_owner_round_robin():
    from dav1d_b200 import shard
    assert [shard.frame_owner(n, 4) for n in range(6)] == [0, 1, 2, 3, 0, 1]


# ------------------------------------------------------------------------------------------ GPU: peer memory
GW, GH, GN = 648, 520, 8


def _worker_gpu(rank, world, port, outdir, n_streams):
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dav1d_b200 import frame, shard, get_lib
        lib = get_lib()
        frames = _frames(10, GW, GH, GN, seed=950)

        def make(S, rows):
            return frame.FrameBuffers(S, band_rows=rows, compact=True)
        pics = shard.decode_gop(frames, make, dist, rank, world, lib, exchange="peer", band_rows=64, n_streams=n_streams)
        np.savez(os.path.join(outdir, "r%d.npz" % rank), **{str(k): v for k, v in pics.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(240)
@pytest.mark.parametrize("world,n_streams", [(2, 1), (2, 2), (4, 1)])
def test_gpu_ranks_shard_a_dependent_gop(tmp_path, world, n_streams):
    """one process per rank; reference rows cross ranks as peer-memory puts gated by flags (NVLink when the ranks have
    their own GPUs); every picture equals the single-process decode and the oracle's chained decode"""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (two processes time-slicing one GPU poll each other for milliseconds per flag)" % world)
    port = 29500 + (os.getpid() + world * 11 + n_streams) % 2000
    mp.spawn(_worker_gpu, args=(world, port, str(tmp_path), n_streams), nprocs=world, join=True)
    frames = _frames(10, GW, GH, GN, seed=950)
    got = _collect(str(tmp_path), world, GN)
    import test_looprestoration as TLR
    exp = oracle_gop(frames)
    for k in range(GN):
        assert TLR.picture_equal(frames[k], got[k], exp[k]), "frame %d differs from the oracle's chained decode" % k


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams", [1, 2])
def test_gpu_single_rank_pipeline(n_streams):
    """world = 1 on the GPU: band pipeline (events between the streams when two frames are in flight) vs the oracle"""
    from dav1d_b200 import frame, shard, get_lib
    import test_looprestoration as TLR
    frames = _frames(8, GW, GH, 6, seed=960)
    exp = oracle_gop(frames)

    def make(S, rows):
        return frame.FrameBuffers(S, band_rows=rows)
    got = shard.decode_gop(frames, make, None, 0, 1, get_lib(), band_rows=128, n_streams=n_streams)
    for k in range(6):
        assert TLR.picture_equal(frames[k], got[k], exp[k]), k


# ------------------------------------------------------------------------------------------ GPU: CUDA graph replay
def _endless_frames(world, n_sets, total):
    """the frame sequence an endless pipeline decodes: rank r's set i holds a fixed synthetic frame, global frame n uses
    set (n // world) % n_sets of rank n % world"""
    base = {(r, i): _frames(8, GW, GH, 1, seed=970 + 16 * r + i)[0] for r in range(world) for i in range(n_sets)}
    return base, [base[(n % world, (n // world) % n_sets)] for n in range(total)]


def _worker_graph(rank, world, port, outdir, n_streams, per_rank):
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dav1d_b200 import frame, shard, get_lib
        lib = get_lib()
        n_sets = 4
        base, _ = _endless_frames(world, n_sets, 0)
        sets = [frame.FrameBuffers(base[(rank, i)], band_rows=64, compact=True) for i in range(n_sets)]
        x = shard.PeerExchange(lib, dist, rank, world, base[(0, 0)]["pic"].nbytes, 2) if world > 1 else None
        pipe = shard.GopPipeline(lib, rank, world, sets, exchange=x, n_streams=n_streams, graphs=True)
        for _ in range(per_rank):
            pipe.submit()
        pipe.sync()
        if world > 1:
            dist.barrier()
        # the last n_sets frames of this rank are still resident
        out = {str((per_rank - n_sets + i) * world + rank): pipe.output(per_rank - n_sets + i) for i in range(n_sets)}
        np.savez(os.path.join(outdir, "g%d.npz" % rank), **out)
        if x is not None:
            x.close()
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(240)
@pytest.mark.parametrize("world,n_streams", [(1, 1), (1, 2), (2, 1), (2, 2)])
def test_gpu_graph_replay_matches_oracle(tmp_path, world, n_streams):
    """from a set's second frame on the band schedule is ONE CUDA graph launch per frame (flag values derived on the device
    from the frame's sequence word): the pictures decoded that way equal the oracle's chained decode of the same stream"""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    per_rank = 10                      # sets are used 2.5 times: eager, captured + replayed, replayed
    port = 29500 + (os.getpid() + world * 13 + n_streams) % 2000
    if world == 1:
        _worker_graph(0, 1, port, str(tmp_path), n_streams, per_rank)
    else:
        mp.spawn(_worker_graph, args=(world, port, str(tmp_path), n_streams, per_rank), nprocs=world, join=True)
    import test_looprestoration as TLR
    _, seq = _endless_frames(world, 4, per_rank * world)
    exp = oracle_gop(seq)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "g%d.npz" % r))
        assert len(z.files) == 4
        for k in z.files:
            assert TLR.picture_equal(seq[int(k)], z[k], exp[int(k)]), "frame %s (rank %d) differs from the oracle's chained decode" % (k, r)
