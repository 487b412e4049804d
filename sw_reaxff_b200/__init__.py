"""sw_reaxff_b200 — B200-native (sm_100a) ReaxFF force-and-charge path.

The product is `librxb200.so` (hand-written CUDA behind the C ABI of include/rxb200.h) plus the C++ host styles in
`host/`.  This Python module is only the ctypes binding used by tests, bench.py and the multi-GPU launcher; it holds no
compute path and NO CPU fallback: importing works anywhere, creating a handle without the built library or without a
CUDA device raises.
"""
from .api import Rxb, RxbError, load_library, LIB_PATH, DATA_DIR  # noqa: F401

__all__ = ["Rxb", "RxbError", "load_library", "LIB_PATH", "DATA_DIR"]
