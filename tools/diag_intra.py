"""GPU diagnostic: intra frames in flight — all-at-once vs staggered submission, device-resident vs host-buffer."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dav1d_b200 import synth, frame, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
G = int(sys.argv[2]) if len(sys.argv) > 2 else 24
lf = int(sys.argv[3]) if len(sys.argv) > 3 else 1
S = [synth.make_intra_frame(np.random.default_rng(1 + k % 4), 8, 1920, 1080) for k in range(4)]
fbs = [frame.FrameBuffers(S[k % 4], run_lf=bool(lf), run_cdef=False, run_lr=False, intra_grid=G, compact=True) for k in range(N)]
streams = [torch.cuda.Stream() for _ in range(N)]
for k in range(N):
    fbs[k].run(streams[k].cuda_stream)
torch.cuda.synchronize()
px = 1920 * 1080


def report(name, dt, frames):
    print("%-44s %.2f ms/frame  %.0f Mpix/s" % (name, dt / frames * 1e3, frames * px / dt / 1e6), flush=True)


# (1) all at once, device resident
t0 = time.perf_counter()
for rep in range(3):
    for k in range(N):
        fbs[k].run(streams[k].cuda_stream)
    torch.cuda.synchronize()
report("all at once (run), N=%d grid=%d lf=%d" % (N, G, lf), time.perf_counter() - t0, 3 * N)
# (2) staggered, device resident: wait for the oldest, resubmit
t0 = time.perf_counter()
M = 4 * N
for i in range(M):
    k = i % N
    streams[k].synchronize()
    fbs[k].run(streams[k].cuda_stream)
torch.cuda.synchronize()
report("staggered (run)", time.perf_counter() - t0, M)
# (3) staggered, host buffers
for k in range(N):
    fbs[k].submit_host()
for k in range(N):
    fbs[k].wait()
t0 = time.perf_counter()
for i in range(M):
    k = i % N
    fbs[k].wait()
    fbs[k].submit_host()
for k in range(N):
    fbs[k].wait()
report("staggered (submit_host)", time.perf_counter() - t0, M)
