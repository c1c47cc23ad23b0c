"""dav1d_b200 — B200-native AV1 block-reconstruction / post-filter back end.

Host-side mirror of dav1d's Dav1dDSPContext surface (dsp.py) over the C ABI in
include/b200av1.h (csrc/*.cu, hand-written sm_100a kernels). See DESIGN.md.
"""
import os as _os
# up to 32 hardware work queues (default 8): the frames in flight and the peer-copy stream of the multi-GPU pipeline each
# get their own, so a stream parked in a flag wait never holds back the stream that feeds the rank it is waiting for
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from ._lib import get_lib, B200Lib, B200Error, ItxBlock  # noqa: F401
from . import levels  # noqa: F401
