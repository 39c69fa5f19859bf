"""buffalo_b200 -- B200-native implementation of kakao/buffalo's matrix-factorisation training hot path
(ALS row solves, BPRMF / WARP negative-sampling SGD) behind buffalo's own Python API.

The compute lives in buffalo_b200/csrc (hand-written sm_100a CUDA behind the C ABI of
include/buffalo_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
