"""pytrimal_b200 -- trimAl's per-alignment statistics hot path on NVIDIA B200.

The product is ``libtrimal_cuda.so`` (hand-written sm_100a kernels behind the
C ABI in ``include/trimal_cuda.h``); this package is the thin host layer that
mirrors the parts of pytrimal's API the path touches.  There is no CPU
fallback: importing works anywhere the library has been built, computing needs
a B200.
"""
from ._lib import (LIB_PATH, NoDeviceError, SymbolError, TrimalCudaError, device_count, load)
from .alignment import Alignment
from .matrix import SimilarityMatrix
from .statistics import (Communicator, DeviceAlignment, cluster_order, gaps_window, get_devices,
                         set_devices, shard_blocks, shard_range, similarity_window,
                         threshold_rule)

__version__ = "0.1.0"

__all__ = [
    "Alignment", "SimilarityMatrix", "DeviceAlignment", "Communicator", "cluster_order", "gaps_window",
    "similarity_window", "threshold_rule", "shard_blocks", "shard_range", "set_devices", "get_devices",
    "device_count", "load", "NoDeviceError", "SymbolError", "TrimalCudaError", "LIB_PATH",
]
