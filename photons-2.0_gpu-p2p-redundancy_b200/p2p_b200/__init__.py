"""p2p_b200 -- Python host-side front-end of the B200 near-field P2P library.

This package is plumbing over the C-ABI in include/p2p_b200.h (ctypes; no torch types cross the
boundary).  It never imports anything from oracle/ and has no CPU fallback: if the CUDA library is
missing or no sm_100 device is present, construction raises.
"""
from .binding import LIB_DIR, P2PContext, P2PError, build_library, device_count, load_library  # noqa: F401
