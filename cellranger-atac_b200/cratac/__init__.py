"""cratac: B200-native COUNT_CUT_SITES + DETECT_PEAKS + GENERATE_PEAK_MATRIX.

Host side of libcratac.so (hand-written sm_100a CUDA behind a C ABI, include/cratac.h).
The product path has no CPU fallback: importing ``cratac._ffi`` and creating a
``Context`` fails loudly when the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
