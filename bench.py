#!/usr/bin/env python3
"""bench.py — throughput of the B200 AV1 reconstruction back end on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic block records.
Prints ONE JSON line (rank 0). See DESIGN.md §Measurement for the byte accounting.

Workloads
  itx8x8   BASELINE config 0: 2^20 inv_txfm_add DCT_DCT 8x8, 8-bit, checkasm-style coefficients
           on an 8192x8192 plane (67.1 Mpx / step). Algorithmic bytes: 256 B / block
           (128 B coefs + 64 B dst read + 64 B dst write).

--impl reference times dav1d's own C functions (oracle/_ref, unmodified reference sources)
on the host cores over a bounded sample of the same records.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ITX_DT = np.dtype([("dst_off", "<u4"), ("coef_off", "<u4"), ("eob", "<i2"), ("txtp", "u1"), ("plane", "u1")])


# ------------------------------------------------------------------------------ workload
def fdct_matrix(n):
    i = np.arange(n)[:, None].astype(np.float64)
    j = np.arange(n)[None, :].astype(np.float64)
    m = np.cos(np.pi * (2 * j + 1) * i / (2.0 * n))
    m[0] *= np.sqrt(0.5)
    return m


def make_itx8x8(seed, n_blocks, plane_w):
    """BASELINE config 0 records (vectorised port of the checkasm generator's distribution:
    random +-255 residual -> float forward DCT x 2.0 -> round; eob uniform over the dc-only /
    full classes; reference tests/checkasm/itx.c:185-242)."""
    rng = np.random.default_rng(seed)
    m = fdct_matrix(8)
    coefs = np.empty((n_blocks, 64), np.int16)
    eobs = np.empty(n_blocks, np.int16)
    import refs
    order = refs.scan_table(1)          # dav1d_scans[TX_8X8]: scan position -> coefficient index
    inv = np.empty(64, np.int32)
    inv[order] = np.arange(64)
    chunk = 1 << 16
    for s in range(0, n_blocks, chunk):
        e = min(n_blocks, s + chunk)
        resid = rng.integers(-255, 256, (e - s, 8, 8)).astype(np.float64)
        out = np.einsum("ij,njk,lk->nil", m, resid.transpose(0, 2, 1), m) * 2.0   # [n][x][y]
        c = np.trunc(out.reshape(e - s, 64) + 0.5).astype(np.int64)
        dc_only = rng.integers(0, 2, e - s) == 0
        eob = np.where(dc_only, 0, rng.integers(1, 63, e - s))
        c[inv[None, :] > eob[:, None]] = 0
        coefs[s:e] = c.astype(np.int16)
        eobs[s:e] = eob
    per_row = plane_w // 8
    blocks = np.zeros(n_blocks, ITX_DT)
    i = np.arange(n_blocks)
    blocks["dst_off"] = (i // per_row) * 8 * plane_w + (i % per_row) * 8
    blocks["coef_off"] = i * 64
    blocks["eob"] = eobs
    blocks["txtp"] = 0
    rows = (n_blocks + per_row - 1) // per_row * 8
    pic = rng.integers(0, 256, (rows, plane_w), dtype=np.uint8)
    return blocks, coefs.reshape(-1), pic


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, False, []

    def run(self):
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    import refs
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    n = 1 << 18                                        # bounded sample of the same workload
    blocks, coefs, pic = make_itx8x8(1, n, 8192)
    lib = refs.ref()
    st = (C.c_int32 * 3)(8192, 8192, 8192)

    def step():
        return lib.refdrv_itx_add_batch(255, 1, blocks.ctypes.data, n, coefs.ctypes.data, pic.ctypes.data, st, 0, ncores)
    for _ in range(args.warmup):
        step()
    t = [step() for _ in range(args.steps)]
    ms = 1e3 * sum(t) / len(t)
    val = n * 64 / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "Mpixels/s", "value": val, "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16->i32", "data": "synthetic",
            "config": {"workload": "itx8x8: inv_txfm_add DCT_DCT 8x8 8-bit (BASELINE config 0), sample of 2^18 of the 2^20 blocks",
                       "l2": "n/a (host)"},
            "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": ncores, "kind": "reference",
                             "sample": "2^18 blocks/step, dav1d C path (HAVE_ASM=0, gcc -O3 -march=x86-64-v3), %d threads" % ncores},
            "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from dav1d_b200 import batch, get_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = get_lib()

    n_blocks, plane_w, nsets = 1 << 20, 8192, 3
    px_per_step = n_blocks * 64
    # every rank processes its own, differently-seeded frame batch (frames shard over GPUs)
    sets = []
    host = None
    for k in range(nsets):
        blocks, coefs, pic = make_itx8x8(1 + rank * 16 + k, n_blocks, plane_w)
        if k == 0:
            host = (blocks, coefs, pic)
        sets.append((torch.from_numpy(blocks.view(np.uint8)).cuda(), torch.from_numpy(coefs).cuda(),
                     torch.from_numpy(pic).cuda()))
    strides = [plane_w] * 3

    def step(i):
        b, c, p = sets[i % nsets]
        batch.itx_add_batch(255, 1, b, c, p, strides)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.b200_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step(i)
        ev[i + 1].record()
    sync_all()
    launches = lib.b200_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * px_per_step / (ms_per_step * 1e-3) / 1e6

    # end to end: host (pinned) buffers through the C ABI, copies inside the timed region
    hb = torch.from_numpy(host[0].view(np.uint8)).pin_memory()
    hc = torch.from_numpy(host[1]).pin_memory()
    hp = torch.from_numpy(host[2].copy()).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        batch.itx_add_batch_host(255, 1, hb, hc, hp, strides)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.itx_add_batch_host(255, 1, hb, hc, hp, strides)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * px_per_step / (float(t.item()) * 1e-3) / 1e6
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = n_blocks * 256
        k_ms = sum(kern_ms) / len(kern_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "itx8x8_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        cpu = cpu_baseline()
        line = {"metric": "Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/i16->i32", "data": "synthetic",
                "config": {"workload": "itx8x8: 2^20 inv_txfm_add DCT_DCT 8x8 8-bit blocks per GPU per step (BASELINE config 0), "
                                       "checkasm-style coefficients, 8192x8192 plane",
                           "l2": "3 rotating input sets (576 MB) > 126 MB L2", "blocks_per_step_per_gpu": n_blocks},
                "roofline": {"bound": "hbm", "kernel": "itx_add_kernel<8,8>", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "Mpixels/s", "h2d_bytes_per_step": int(hb.numel() + hc.numel() * 2 + hp.numel()),
                        "d2h_bytes_per_step": int(hp.numel())},
                "gpu_launches": int(launches), "clocks": sampler.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline():
    """dav1d's C path (oracle/_ref) when shipped, else the oracle port, on a bounded sample."""
    import refs
    n = 1 << 18
    blocks, coefs, pic = make_itx8x8(1, n, 8192)
    st = (C.c_int32 * 3)(8192, 8192, 8192)
    ncores = os.cpu_count() or 1
    if refs.have_ref():
        lib = refs.ref()
        lib.refdrv_itx_add_batch(255, 1, blocks.ctypes.data, n, coefs.ctypes.data, pic.ctypes.data, st, 0, ncores)
        reps, tot = 0, 0.0
        while tot < 3.0 and reps < 200:
            tot += lib.refdrv_itx_add_batch(255, 1, blocks.ctypes.data, n, coefs.ctypes.data, pic.ctypes.data, st, 0, ncores)
            reps += 1
        t1 = lib.refdrv_itx_add_batch(255, 1, blocks.ctypes.data, n, coefs.ctypes.data, pic.ctypes.data, st, 0, 1)
        return {"value": reps * n * 64 / tot / 1e6, "unit": "Mpixels/s", "cores": ncores, "kind": "reference",
                "sample": "2^18 of the 2^20 blocks x %d reps, dav1d C path HAVE_ASM=0 (no nasm in image), %d threads" % (reps, ncores),
                "single_core_value": n * 64 / t1 / 1e6}
    o = refs.oracle()
    t0 = time.perf_counter()
    o.oracle_itx_add_batch(255, 1, blocks.ctypes.data, n, coefs.ctypes.data, pic.ctypes.data, st, 0)
    dt = time.perf_counter() - t0
    return {"value": n * 64 / dt / 1e6, "unit": "Mpixels/s", "cores": 1, "kind": "port", "sample": "2^18 blocks, oracle port, 1 thread"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
