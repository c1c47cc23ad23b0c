"""GPU diagnostic: where does the end-to-end frame time go (uploads / job / downloads), dense vs compact upload."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctypes as C
from dav1d_b200 import synth, frame, _lib

lib = _lib.get_lib()
bpc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = synth.make_inter_frame(np.random.default_rng(1), bpc, 3840, 2160, film_grain=bpc > 8)
for compact in (False, True):
    fbs = [frame.FrameBuffers(S, compact=compact) for _ in range(3)]
    for fb in fbs:
        fb.submit_host(); fb.wait()
    fb = fbs[0]
    st = fb._own_stream[1]
    ts = torch.cuda.Stream()
    def ev():
        return torch.cuda.Event(enable_timing=True)
    # serial pieces on one stream
    with torch.cuda.stream(fb._own_stream[0]):
        e = [ev() for _ in range(4)]
        for rep in range(3):
            e[0].record()
            for x in fb._ups:
                lib.b200_frame_submit_host  # noqa
            t0 = time.perf_counter()
            lib.check(lib.b200_frame_submit_host(C.byref(fb.job), fb._ups, len(fb._ups), None, 0, st), "up+job")
            t1 = time.perf_counter()
            e[1].record()
            lib.check(lib.b200_frame_submit_host(C.byref(_lib.FrameJob()), None, 0, fb._downs, 1, st), "down") if False else None
            e[2].record()
            torch.cuda.synchronize()
        print("compact=%s  uploads=%d (%.1f MB)  H2D+job %.3f ms (host submit call %.3f ms)" %
              (compact, len(fb._ups), fb.h2d_bytes / 1e6, e[0].elapsed_time(e[1]), (t1 - t0) * 1e3))
    # job only
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record(); fb.run(); e1.record(); torch.cuda.synchronize()
    print("   job only (records resident) %.3f ms" % e0.elapsed_time(e1))
    # pipelined, 3 frames in flight
    for n in (30,):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            if i >= 3:
                fbs[i % 3].wait()
            fbs[i % 3].submit_host()
        for f in fbs:
            f.wait()
        dt = (time.perf_counter() - t0) / n
        print("   pipelined x3: %.3f ms/frame -> %.0f Mpix/s" % (dt * 1e3, 3840 * 2160 / dt / 1e6))
    # single stream, sync each frame
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        fbs[0].submit_host(); fbs[0].wait()
    dt = (time.perf_counter() - t0) / 10
    print("   serial (1 in flight): %.3f ms/frame" % (dt * 1e3))
