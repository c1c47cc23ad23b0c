"""Level-2 (batched, device-resident) entry points over torch CUDA tensors.

torch is plumbing only: device memory (tensor.data_ptr()) and the current CUDA stream are
handed to the C ABI (include/b200av1.h); the kernels are ours.
"""
import ctypes as C
import numpy as np

from ._lib import get_lib, ItxBlock

ITX_BLOCK_DTYPE = np.dtype([("dst_off", "<u4"), ("coef_off", "<u4"), ("eob", "<i2"),
                            ("txtp", "u1"), ("plane", "u1")])
assert ITX_BLOCK_DTYPE.itemsize == C.sizeof(ItxBlock) == 12


def _strides3(stride_px):
    s = list(stride_px) + [stride_px[-1]] * (3 - len(stride_px))
    return (C.c_int32 * 3)(*s)


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return int(getattr(stream, "cuda_stream", stream))


def itx_add_batch(bitdepth_max, tx, blocks, coef, pic, stride_px, zero_coefs=False, stream=None, lib=None):
    """blocks: uint8 CUDA tensor holding n x B200ItxBlock; coef/pic: CUDA tensors. Asynchronous."""
    lib = lib or get_lib()
    n = blocks.numel() * blocks.element_size() // ITX_BLOCK_DTYPE.itemsize
    rc = lib.b200_itx_add_batch(bitdepth_max, tx, blocks.data_ptr(), n, coef.data_ptr(), pic.data_ptr(),
                                _strides3(stride_px), int(zero_coefs), _stream_ptr(stream))
    lib.check(rc, "b200_itx_add_batch")


def itx_add_batch_host(bitdepth_max, tx, blocks, coef, pic, stride_px, zero_coefs=False, lib=None):
    """Host-buffer form (numpy arrays or pinned torch CPU tensors); synchronous; pic updated in place."""
    lib = lib or get_lib()

    def ptr_bytes(a):
        if isinstance(a, np.ndarray):
            return a.ctypes.data, a.nbytes
        return a.data_ptr(), a.numel() * a.element_size()
    bp, bb = ptr_bytes(blocks)
    cp, cb = ptr_bytes(coef)
    pp, pb = ptr_bytes(pic)
    rc = lib.b200_itx_add_batch_host(bitdepth_max, tx, bp, bb // ITX_BLOCK_DTYPE.itemsize, cp, cb, pp, pb,
                                     _strides3(stride_px), int(zero_coefs))
    lib.check(rc, "b200_itx_add_batch_host")
