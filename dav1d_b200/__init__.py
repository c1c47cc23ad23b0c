"""dav1d_b200 — B200-native AV1 block-reconstruction / post-filter back end.

Host-side mirror of dav1d's Dav1dDSPContext surface (dsp.py) over the C ABI in
include/b200av1.h (csrc/*.cu, hand-written sm_100a kernels). See DESIGN.md.
"""
from ._lib import get_lib, B200Lib, B200Error, ItxBlock  # noqa: F401
from . import levels  # noqa: F401
