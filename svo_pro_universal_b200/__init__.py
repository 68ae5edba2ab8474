"""B200-native direct front-end hot path of SVO Pro (pyramid + FAST, sparse image alignment, patch matching, depth filter).

The product is libsvo_cuda.so (hand-written sm_100a kernels behind the C ABI of include/svo_cuda.h) plus the C++ facades in
host/. This Python package only binds that C ABI (capi), generates synthetic inputs (synth) and packs batches (batch).
"""
from . import capi, synth, batch  # noqa: F401

__all__ = ["capi", "synth", "batch"]
