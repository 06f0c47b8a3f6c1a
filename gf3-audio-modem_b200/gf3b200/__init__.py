"""gf3b200 -- host side of the B200-native GF3 OFDM physical layer.

    from gf3b200 import Phy          # device-level API (torch tensors)
    import OFDM                      # drop-in for the reference's OFDM.py (numpy in / out)

Everything numerical runs in libgf3b200.so (hand-written sm_100a kernels) through ctypes; there
is no CPU fallback.
"""
from ._lib import Gf3Error, Gf3Params, launch_count, load  # noqa: F401
from .phy import MODES, Phy, default_known_sequence, qpsk_points  # noqa: F401

__all__ = ["Phy", "MODES", "Gf3Error", "Gf3Params", "load", "launch_count", "default_known_sequence", "qpsk_points"]
