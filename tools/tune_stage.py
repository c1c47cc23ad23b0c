"""GPU tuning aid: time single stages of the 4K frame job for several builds of the library
(dav1d_b200.build.build_variant). usage: python tools/tune_stage.py [bpc] lib1.so lib2.so ..."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from dav1d_b200 import synth, frame, _lib

bpc = int(sys.argv[1])
paths = sys.argv[2:]
Ss = [synth.make_inter_frame(np.random.default_rng(1 + k), bpc, 3840, 2160, film_grain=bpc > 8) for k in range(3)]
for path in paths:
    lib = _lib.B200Lib(path)
    fbs = [frame.FrameBuffers(S, lib=lib) for S in Ss]
    st = torch.cuda.current_stream().cuda_stream
    stages = [("pred", lambda j, bd: lib.b200_mc_batch(bd, C.byref(j.mc), j.d_pred, j.n_pred, st)),
              ("comp", lambda j, bd: lib.b200_mc_comp_batch(bd, C.byref(j.mc), j.d_comp, j.n_comp, st)),
              ("itx", lambda j, bd: lib.b200_itx_add_frame(bd, j.d_itx, j.n_itx, j.d_coef, j.mc.dst, j.itx_stride, 0, st)),
              ("deblock", lambda j, bd: lib.b200_lf_frame(bd, C.byref(j.lf), st)),
              ("cdef", lambda j, bd: lib.b200_cdef_frame(bd, C.byref(j.cdef), st)),
              ("lr", lambda j, bd: lib.b200_lr_frame(bd, C.byref(j.lr), st))]
    if bpc > 8:
        stages.append(("fg", lambda j, bd: lib.b200_fg_apply(bd, C.byref(j.fg), st)))
    acc = {n: [] for n, _ in stages}
    for rep in range(13):
        j = fbs[rep % 3].job
        for n, fn in stages:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(j, j.bitdepth_max); e1.record()
            torch.cuda.synchronize()
            if rep >= 1:
                acc[n].append(e0.elapsed_time(e1))
    print(os.path.basename(path), " ".join("%s=%.1fus" % (n, 1e3 * float(np.median(v))) for n, v in acc.items()), flush=True)
    del fbs
