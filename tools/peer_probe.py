"""GPU probe of the exchange primitives (include/b200av1.h: b200_ipc_*, b200_copy_async, b200_flag_*): two processes (one GPU
each when there are two, else sharing cuda:0) map each other's arena, ping-pong flags and put data. Prints what works.
    python tools/peer_probe.py [world]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def say(rank, *a):
    print("[rank %d %.2f]" % (rank, time.time() % 1000), *a, flush=True)


def worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dav1d_b200 import get_lib
    lib = get_lib()
    say(rank, "devices", ndev, "flag kernels forced" if os.environ.get("B200_FLAG_KERNELS") else "cuStreamWaitValue32 if available")
    n = 1 << 20
    arena = lib.b200_dev_alloc(4096 + n)
    lib.check(lib.b200_dev_memset(arena, 0, 4096 + n, None), "memset"); lib.check(lib.b200_frame_wait(None), "sync")
    h = (C.c_uint8 * 64)()
    lib.check(lib.b200_ipc_export(arena, h), "export")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(h))
    peer_rank = (rank + 1) % world
    peer = lib.b200_ipc_open((C.c_uint8 * 64).from_buffer_copy(handles[peer_rank]))
    assert peer, lib.b200_last_error()
    say(rank, "ipc open ok")
    dist.barrier()
    st = lib.b200_stream_create()
    src = torch.full((n,), rank + 1, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    # put my data into the peer's arena, then raise its flag; wait for my own flag; check what landed
    t0 = time.time()
    lib.check(lib.b200_copy_async(peer + 4096, src.data_ptr(), n, st), "copy")
    lib.check(lib.b200_flag_signal(peer + 64, 1, st), "signal")
    lib.check(lib.b200_flag_wait_geq(arena + 64, 1, st), "wait")
    out = np.zeros(n, np.uint8)
    lib.check(lib.b200_copy_async(out.ctypes.data, arena + 4096, n, st), "d2h")
    lib.check(lib.b200_frame_wait(st), "sync")
    say(rank, "put + flag round ok: landed value", int(out[0]), int(out[-1]), "expected", (rank - 1) % world + 1, "%.1f ms" % ((time.time() - t0) * 1e3))
    assert (out == (rank - 1) % world + 1).all()
    # ping-pong latency: rank 0 signals k, rank 1 waits k and signals back
    dist.barrier()
    t0 = time.time()
    K = 200
    for k in range(2, K + 2):
        if rank == 0:
            lib.check(lib.b200_flag_signal(peer + 128, k, st), "signal")
            lib.check(lib.b200_flag_wait_geq(arena + 128, k, st), "wait")
        elif rank == 1:
            lib.check(lib.b200_flag_wait_geq(arena + 128, k, st), "wait")
            lib.check(lib.b200_flag_signal(peer + 128, k, st), "signal")
    lib.check(lib.b200_frame_wait(st), "sync")
    if rank < 2 and world == 2:
        say(rank, "flag ping-pong: %.1f us per round trip" % ((time.time() - t0) / K * 1e6))
    # peer copy bandwidth
    big = 256 << 20
    a2 = lib.b200_dev_alloc(big)
    h2 = (C.c_uint8 * 64)(); lib.check(lib.b200_ipc_export(a2, h2), "export")
    hs = [None] * world; dist.all_gather_object(hs, bytes(h2))
    p2 = lib.b200_ipc_open((C.c_uint8 * 64).from_buffer_copy(hs[peer_rank]))
    srcb = torch.zeros(big, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize(); dist.barrier()
    for rep in range(2):
        t0 = time.time()
        lib.check(lib.b200_copy_async(p2, srcb.data_ptr(), big, st), "copy"); lib.check(lib.b200_frame_wait(st), "sync")
        dt = time.time() - t0
    say(rank, "peer copy %.1f GB/s (256 MiB)" % (big / dt / 1e9))
    for sz in (64 << 10, 1 << 20):
        t0 = time.time()
        for _ in range(50):
            lib.check(lib.b200_copy_async(p2, srcb.data_ptr(), sz, st), "copy")
        lib.check(lib.b200_frame_wait(st), "sync")
        say(rank, "peer copy of %d KiB: %.1f us each" % (sz >> 10, (time.time() - t0) / 50 * 1e6))
    dist.barrier()
    lib.b200_ipc_close(peer); lib.b200_ipc_close(p2)
    dist.destroy_process_group()
    say(rank, "done")


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mp.spawn(worker, args=(world, 29400 + os.getpid() % 500), nprocs=world, join=True)
