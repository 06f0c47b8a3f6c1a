"""Import shim for the UNMODIFIED reference module (/root/reference/OFDM.py).

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference).  Used by oracle/make_golden.py to generate tests/golden/*.npz and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent) to pin the numpy
oracle against the real implementation.

Recipe:
  * stub the modules OFDM.py imports at top level but that are not installed here
    (matplotlib, sounddevice, IPython, pyldpc)            -- OFDM.py:4,7-10
  * chdir into a scratch directory holding lower-case symlinks, because the reference opens
    "handouts/...", "input_files/...", "output_files/..." -- OFDM.py:99,757,793
  * attribute-patch a constructed object to any (N, CP, lo, hi, P, L)
"""
import contextlib
import io
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("GF3_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "OFDM.py"))


_scratch = None


def scratch_dir():
    global _scratch
    if _scratch is None:
        d = tempfile.mkdtemp(prefix="gf3ref_")
        os.symlink(os.path.join(REF_ROOT, "Handouts"), os.path.join(d, "handouts"))
        os.symlink(os.path.join(REF_ROOT, "input_Files"), os.path.join(d, "input_files"))
        os.symlink(os.path.join(REF_ROOT, "received_signals"), os.path.join(d, "received_signals"))
        os.symlink(os.path.join(REF_ROOT, "sound_files"), os.path.join(d, "sound_files"))
        os.makedirs(os.path.join(d, "output_files"))
        _scratch = d
    return _scratch


def load():
    """Return the reference OFDM module (imported once, cwd switched to the scratch dir)."""
    if "OFDM" in sys.modules and getattr(sys.modules["OFDM"], "__gf3_reference__", False):
        return sys.modules["OFDM"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "sounddevice", "IPython", "IPython.display", "pyldpc"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["sounddevice"].default = types.SimpleNamespace(channels=1)
    sys.modules["IPython.display"].Audio = object
    sys.modules["IPython"].display = sys.modules["IPython.display"]
    os.chdir(scratch_dir())
    saved = sys.modules.pop("OFDM", None)
    sys.path.insert(0, REF_ROOT)
    try:
        import OFDM  # noqa
    finally:
        sys.path.remove(REF_ROOT)
    OFDM.__gf3_reference__ = True
    if saved is not None:  # keep the reference reachable under a private name too
        sys.modules["OFDM_reference"] = OFDM
    return OFDM


def make(cls_name, mode="A2", encoding="XOR", no_pilots=20, packet_length=180,
         N=None, cp=None, lo=None, hi=None):
    """Construct reference class `cls_name` and attribute-patch it to (N, cp, lo, hi).

    Every reference method reads self.* dynamically, so overwriting the
    derived attributes of CamG.__init__ (OFDM.py:27-95) re-parameterises the object.
    """
    import numpy as np
    ref = load()
    os.chdir(scratch_dir())
    obj = getattr(ref, cls_name)(mode, encoding=encoding, no_pilots=no_pilots, packet_length=packet_length)
    if N is not None:
        obj.ofdm_symbol_size = N
        obj.K = N // 2 - 1
        obj.L = obj.K + 1
    if cp is not None:
        obj.cp_length = cp
    if lo is not None:
        obj.lowest_bin = lo
    if hi is not None:
        obj.highest_bin = hi
    obj.carriers = np.arange(1, obj.K + 1)
    obj.data_carriers = np.arange(obj.lowest_bin, obj.highest_bin)
    obj.data_carriers_per_symbol = len(obj.data_carriers)
    obj.unused_carriers = np.delete(obj.carriers, (obj.data_carriers - 1))
    obj.chirp_length = 5 * (obj.ofdm_symbol_size + obj.cp_length)
    obj.data_bits_per_symbol = obj.data_carriers_per_symbol * obj.mu
    obj.bits_per_symbol = obj.K * obj.mu
    return obj


@contextlib.contextmanager
def quiet():
    """Swallow the reference's print banners."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf
