#!/usr/bin/env python3
"""Times the HOST side of the hooked decoder (dav1d's front end + the record emitters + frame completion) with the back end that
does nothing (tools/null_backend.c) next to stock dav1d on the same stream and thread count. No GPU involved: this is the part of a
stream decode the device cannot speed up, and what the emitters cost on top of dav1d's own parsing.
usage: tools/host_side.py [threads] [workload ...]   -> markdown table on stdout"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                     # noqa: E402
import refs                      # noqa: E402
import streamgen                 # noqa: E402
from dav1d_b200 import obu, stream   # noqa: E402


def main():
    nthr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    wls = sys.argv[2:] or ["stream1080p8_inter", "stream1080p8_sparse", "stream4k8_inter", "stream4k8_sparse"]
    null_so = "/tmp/libb200_null.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", null_so, os.path.join(ROOT, "tools", "null_backend.c")], check=True)
    dec = stream.HookedDecoder(backend=null_so)
    ref = C.CDLL(refs.REF_SO)
    print("| workload | stream | hooked decoder, host side only | stock dav1d (whole decode) | host side / stock |\n|---|---|---|---|---|")
    for wl in wls:
        W = bench.STREAM_WORKLOADS[wl]
        gen = (lambda *a, **k: obu.inter_stream(*a, motion_modes=2, **k)) if W.get("inter") else obu.intra_stream
        build = lambda: gen(100, W["W"], W["H"], n_frames=W["frames"], bpc=W["bpc"], log2_cols=W["log2_cols"], log2_rows=W["log2_rows"])
        tus = streamgen.generate(build, seed=100, check=False, **W["gen"])[0] if W.get("gen") else build()
        stream.decode_stream.capacity = (W["W"] * W["H"] * 3 // 2) * (2 if W["bpc"] > 8 else 1) * W["frames"] + (1 << 20)
        mfd = min(8, W["frames"], nthr)
        best = [1e9, 1e9]
        for rep in range(7):
            for k, dll in enumerate((dec.dll, ref)):
                if k and rep >= 3:
                    continue
                t0 = time.perf_counter()
                r, _, _ = stream.decode_stream(dll, tus, n_threads=nthr, max_frame_delay=mfd)
                assert r == W["frames"], r
                best[k] = min(best[k], (time.perf_counter() - t0) * 1e3 / W["frames"])
        dec.stats(reset=True)
        print("| `%s` | %d frames %dx%d, %.1f KB per frame | %.1f ms per frame | %.1f ms per frame | %.0f %% |"
              % (wl, W["frames"], W["W"], W["H"], len(b"".join(tus)) / W["frames"] / 1e3, best[0], best[1], 100 * best[0] / best[1]), flush=True)


if __name__ == "__main__":
    main()
