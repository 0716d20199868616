"""dpdist_b200: B200-native implementation of DPDist's hot path (3DmFV -> local patches ->
implicit distance MLP) behind the reference's own Python API.

    from dpdist_b200 import dpdist_and_aue as MODEL      # models/dpdist_and_aue.py
    from dpdist_b200 import dpdist_util                  # utils/dpdist_util.py
    from dpdist_b200 import tf_util                      # utils/tf_util.py (variables only)

All compute runs in libdpdist_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/dpdist_b200.h).  There is no CPU fallback.
"""
from . import _lib, tf_util, dpdist_util, dpdist_and_aue  # noqa: F401

__all__ = ["_lib", "tf_util", "dpdist_util", "dpdist_and_aue"]
