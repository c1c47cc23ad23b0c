#!/usr/bin/env python3
"""Fuzz the multi-rank decode of a dependent group of pictures on the CPU: 2 - 4 gloo ranks drive the host-emulator build, frame n
on rank n mod world predicts from the restored pictures of frames n-1 / n-2 that other ranks own, bands of 64 - 192 rows, random
frame sizes / bit depths / GOP lengths (incl. GOPs shorter than the rank count); every picture must equal the oracle's chained
decode (tests/test_multigpu.py). usage: tools/fuzz_gop.py [seed] [n_configs]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist, torch.multiprocessing as mp
import refs, test_multigpu as TM

def worker(rank, world, port, outdir, bpc, w, h, n, seed, rows):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ["MASTER_ADDR"]="127.0.0.1"; os.environ["MASTER_PORT"]=str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pics = TM._decode_emu(rank, world, TM._frames(bpc, w, h, n, seed), band_rows=rows)
        np.savez(os.path.join(outdir, "r%d.npz" % rank), **{str(k): v for k, v in pics.items()})
    finally:
        dist.destroy_process_group()

if __name__ == "__main__":
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    bad = 0; t0 = time.time(); n_cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    for it in range(n_cfg):
        world = int(rng.choice([2, 3, 4])); bpc = int(rng.choice([8, 10])); w = int(rng.integers(12, 40)) * 8; h = int(rng.integers(20, 70)) * 8
        n = int(rng.integers(2, 3 * world + 3)); rows = 64 * int(rng.integers(1, 4)); seed = int(rng.integers(0, 10000))
        frames = TM._frames(bpc, w, h, n, seed)
        exp = TM.oracle_gop(frames)
        with tempfile.TemporaryDirectory() as d:
            mp.spawn(worker, args=(world, 29600 + it, d, bpc, w, h, n, seed, rows), nprocs=world, join=True)
            got = TM._collect(d, world, n)
        ok = all(np.array_equal(a, b) for a, b in zip(exp, got))
        bad += not ok
        print("world %d bpc %d %dx%d frames %d band %d: %s" % (world, bpc, w, h, n, rows, "ok" if ok else "MISMATCH"), flush=True)
    print("gop fuzz: %d configs, %d bad, %.0f s" % (n_cfg, bad, time.time() - t0))
