"""ctypes binding of libgf3b200.so (include/gf3_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at load time,
and every compute entry point fails with Gf3Error when no CUDA device is present.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# GF3_LIB_PATH selects an alternate build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("GF3_LIB_PATH") or os.path.join(os.path.dirname(HERE), "lib", "libgf3b200.so")

GF3_OK, GF3_ERR_INVALID, GF3_ERR_CUDA, GF3_ERR_NODEVICE = 0, -1, -2, -3


class Gf3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgf3b200 error %d: %s" % (code, msg))
        self.code = code


class Gf3Params(ctypes.Structure):
    """struct gf3_params -- parameter contract of CamG.__init__ (OFDM.py:18-101)."""
    _fields_ = [
        ("N", c_int32), ("cp", c_int32), ("lo", c_int32), ("hi", c_int32),
        ("n_pilots", c_int32), ("packet_len", c_int32), ("fit_lo", c_int32), ("fit_hi", c_int32),
        ("chirp_len", c_int32),
        ("fs", c_float), ("f0", c_float), ("f1", c_float), ("thresh", c_float),
        ("tx_gain", c_float), ("chirp_gain", c_float),
    ]


# name -> (restype, argtypes); must list every function include/gf3_b200.h declares
SIGNATURES = {
    "gf3_abi_version": (c_int, []),
    "gf3_last_error": (c_char_p, []),
    "gf3_device_count": (c_int, []),
    "gf3_params_default": (c_int, [POINTER(Gf3Params), c_int, c_int, c_int, c_int, c_int, c_int]),
    "gf3_plan_create": (c_int, [POINTER(Gf3Params), POINTER(c_void_p)]),
    "gf3_plan_destroy": (c_int, [c_void_p]),
    "gf3_plan_params": (c_int, [c_void_p, POINTER(Gf3Params)]),
    "gf3_launch_count": (c_int64, []),
    "gf3_rx_estimate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gf3_rx_demod": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_rx_receive": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_rx_receive_pcm": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_rx_receive_is_fused": (c_int, [c_void_p]),
    "gf3_rx_known_channel": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_rx_spectrum": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_sync_chirp": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gf3_xcorr_work_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "gf3_xcorr": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gf3_peak_pick_work_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "gf3_peak_pick": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "gf3_tx_modulate": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "gf3_tx_encode_modulate": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "gf3_tx_ifft": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_cdiv": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "gf3_eq_estimate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gf3_eq_apply": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gf3_demap": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gf3_tx_frame": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "gf3_channel_sim": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_uint64, c_void_p, c_int64, c_void_p]),
    "gf3_channel_sim_ids": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_uint64, c_void_p, c_int64, c_void_p]),
    "gf3_random_bytes": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_uint64, c_void_p]),
    "gf3_sync_work_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "gf3_sync_streams": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "gf3_sync_detect": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "gf3_schmidlcox": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "gf3_peaks_to_offsets": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "gf3_ber_count": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gf3_pcm_to_f32": (c_int, [c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libgf3b200.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libgf3b200.so not found at %s -- build it with `python gf3-audio-modem_b200/build.py` "
            "(the GF3 B200 physical layer has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.gf3_abi_version() != 1:
        raise RuntimeError("libgf3b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != GF3_OK:
        raise Gf3Error(rc, load().gf3_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().gf3_launch_count())
