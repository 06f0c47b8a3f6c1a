"""Device-level GF3 physical layer: torch tensors in HBM in, torch tensors out.

`Phy` wraps one gf3_plan (include/gf3_b200.h) and exposes the fused kernels batch-wise; the
numpy drop-in surface of the reference (OFDM.py: CamG / transmitter / receiver) sits on top of it
in ../OFDM.py.  PyTorch is used only for device memory, streams and (in bench.py / sweep.py)
torch.distributed -- every arithmetic step of the path runs in libgf3b200.so.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import Gf3Params, check

_HERE = os.path.dirname(os.path.abspath(__file__))
KNOWN_SEQUENCE_FILE = os.path.join(_HERE, "known_sequence_4096.txt")

# OFDM.py:30-40
MODES = {
    "A1": (224, (1, 2047)), "A2": (224, (100, 1500)), "A3": (224, (100, 1000)),
    "B1": (704, (1, 2047)), "B2": (704, (100, 1500)), "B3": (704, (100, 1000)),
    "C1": (1184, (1, 2047)), "C2": (1184, (100, 1500)), "C3": (1184, (100, 1000)),
}


def default_known_sequence():
    """The CamG-standard known bit sequence: first 4096 characters of Handouts/random_bits.txt
    (OFDM.py:99-101), shipped with the package so the GPU box does not need the reference tree."""
    raw = np.fromfile(KNOWN_SEQUENCE_FILE, dtype=np.uint8)
    return (raw - ord("0")).astype(np.int64)


def qpsk_points(bits2):
    """Gray QPSK table of OFDM.py:72-77 as an array expression (host-side formatting only)."""
    b = np.asarray(bits2).astype(np.int64)
    return ((1 - 2 * b[..., 1]) + 1j * (1 - 2 * b[..., 0])) / np.sqrt(2)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


_STREAM = object()      # placeholder for "the current CUDA stream of the plan's device" in Phy._call


class Phy:
    """One parameter set of the modem on one CUDA device."""

    def __init__(self, N=4096, cp=224, lo=100, hi=1500, n_pilots=20, packet_len=180,
                 known_sequence=None, device=None, fit_lo=None, fit_hi=None, chirp_len=None,
                 thresh=None, fs=None, f0=None, f1=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available() or self.lib.gf3_device_count() == 0:
            raise _lib.Gf3Error(_lib.GF3_ERR_NODEVICE,
                                "no CUDA device: the GF3 B200 physical layer has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        p = Gf3Params()
        check(self._call("gf3_params_default", ctypes.byref(p), N, cp, lo, hi, n_pilots, packet_len))
        for name, val in (("fit_lo", fit_lo), ("fit_hi", fit_hi), ("chirp_len", chirp_len),
                          ("thresh", thresh), ("fs", fs), ("f0", f0), ("f1", f1)):
            if val is not None:
                setattr(p, name, val)
        self.params = p
        self.N, self.cp, self.lo, self.hi = N, cp, lo, hi
        self.P, self.L = n_pilots, packet_len
        self.K = N // 2 - 1
        self.Nd = hi - lo
        self.symlen = N + cp
        self.chirp_len = int(p.chirp_len)
        self.pkt_samples = (2 * n_pilots + packet_len) * self.symlen
        self.bits_per_packet = 2 * self.Nd * packet_len
        # row stride of packed-bit buffers: whole 32-bit words, 16-byte multiple for vector access
        self.bits_stride = ((self.bits_per_packet + 31) // 32 * 4 + 15) // 16 * 16
        self._plan = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self._call("gf3_plan_create", ctypes.byref(p), ctypes.byref(self._plan)))
        ks = default_known_sequence() if known_sequence is None else np.asarray(known_sequence).astype(np.int64)
        if len(ks) < 2 * self.K:
            raise ValueError("known_sequence needs at least 2K = %d bits" % (2 * self.K))
        self.known_sequence = ks
        known = qpsk_points(ks[: 2 * self.K].reshape(self.K, 2)).astype(np.complex64)     # OFDM.py:429
        self.known = torch.from_numpy(known).to(self.device)
        # does gf3_rx_receive run the channel estimate inside the data-symbol launch for this geometry?
        self.fused_receive = bool(self.lib.gf3_rx_receive_is_fused(self._plan))
        # float32 raw streams: staged input for the receive chain?  (GF3_STREAMS_STAGED=0/1 overrides the default)
        self.staged_streams = os.environ.get("GF3_STREAMS_STAGED", "0") == "1"
        kb = ks[: 2 * self.Nd].reshape(self.Nd, 2)
        self.xor2 = torch.from_numpy(((kb[:, 0] << 1) | kb[:, 1]).astype(np.uint8)).to(self.device)   # OFDM.py:542

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self.lib.gf3_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _call(self, name, *args):
        """Call a library entry point with the plan's device current (kernels launch on the device that
        owns the plan's tables and the caller's tensors, whatever torch's current device is) and
        _STREAM replaced by that device's current stream."""
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            return getattr(self.lib, name)(*[st if a is _STREAM else a for a in args])

    def _f32(self, t):
        assert t.is_cuda and t.dtype == torch.float32 and t.stride(-1) == 1, "need a float32 CUDA tensor with unit sample stride"
        return t

    def _offsets(self, pkt_offset, n):
        if pkt_offset is None:
            return None
        assert pkt_offset.is_cuda and pkt_offset.dtype == torch.int64 and pkt_offset.numel() == n
        return pkt_offset.contiguous()

    # ------------------------------------------------------------------ receive chain
    def rx_estimate(self, samples, n_packets, pkt_offset=None):
        """-> Hs, He complex64 [n_packets, K], slope float64 [n_packets]  (OFDM.py:429-462)."""
        self._f32(samples)
        off = self._offsets(pkt_offset, n_packets)
        Hs = torch.empty((n_packets, self.K), dtype=torch.complex64, device=self.device)
        He = torch.empty_like(Hs)
        slope = torch.empty((n_packets,), dtype=torch.float64, device=self.device)
        check(self._call("gf3_rx_estimate", self._plan, _ptr(samples), _ptr(off), n_packets, _ptr(self.known),
                                       _ptr(Hs), _ptr(He), _ptr(slope), _STREAM))
        return Hs, He, slope

    def rx_demod(self, samples, n_packets, Hs, He, slope, pkt_offset=None, xor=True, want_eq=False,
                 out=None):
        """-> packed bits uint8 [n_packets, bits_stride] (and eq complex64 [n_packets, L, K]).
        Fused CP strip + FFT + equalise + demap + XOR decode (OFDM.py:407-418,593,466-505,541-544)."""
        self._f32(samples)
        off = self._offsets(pkt_offset, n_packets)
        bits = out if out is not None else torch.empty((n_packets, self.bits_stride), dtype=torch.uint8, device=self.device)
        eq = torch.empty((n_packets, self.L, self.K), dtype=torch.complex64, device=self.device) if want_eq else None
        check(self._call("gf3_rx_demod", self._plan, _ptr(samples), _ptr(off), n_packets, _ptr(Hs), _ptr(He), _ptr(slope),
                                    _ptr(self.xor2) if xor else None, _ptr(bits), self.bits_stride, _ptr(eq), _STREAM))
        return (bits, eq) if want_eq else bits

    def rx_known_channel(self, samples, n_packets, Hinv, pkt_offset=None, xor=False, want_eq=False):
        """Known-channel receiver (Weekend Challenge.ipynb:162-226): Y * Hinv on bins 1..K, demap."""
        self._f32(samples)
        off = self._offsets(pkt_offset, n_packets)
        assert Hinv.is_cuda and Hinv.dtype == torch.complex64 and Hinv.numel() == self.K
        bits = torch.empty((n_packets, self.bits_stride), dtype=torch.uint8, device=self.device)
        eq = torch.empty((n_packets, self.L, self.K), dtype=torch.complex64, device=self.device) if want_eq else None
        check(self._call("gf3_rx_known_channel", self._plan, _ptr(samples), _ptr(off), n_packets, _ptr(Hinv.contiguous()),
                                            _ptr(self.xor2) if xor else None, _ptr(bits), self.bits_stride, _ptr(eq), _STREAM))
        return (bits, eq) if want_eq else bits

    def spectrum(self, samples, n_symbols, sym_offset=None):
        """FFT bins 1..K of n_symbols symbols (CP stripped): complex64 [n_symbols, K] (OFDM.py:407-416,593)."""
        self._f32(samples)
        off = self._offsets(sym_offset, n_symbols)
        out = torch.empty((n_symbols, self.K), dtype=torch.complex64, device=self.device)
        check(self._call("gf3_rx_spectrum", self._plan, _ptr(samples), _ptr(off), n_symbols, _ptr(out), _STREAM))
        return out

    def unpack_bits(self, packed, n_packets=None):
        """Host-side: packed [n_packets, bits_stride] -> int64 bit vector (np.unpackbits order)."""
        a = packed.cpu().numpy() if isinstance(packed, torch.Tensor) else np.asarray(packed)
        bits = np.unpackbits(a, axis=1)[:, : self.bits_per_packet]
        return bits.reshape(-1).astype(np.int64)

    # ------------------------------------------------------------------ synchronisation
    def sync_chirp(self):
        out = torch.empty((self.chirp_len,), dtype=torch.float32, device=self.device)
        check(self._call("gf3_sync_chirp", self._plan, _ptr(out), _STREAM))
        return out

    def xcorr(self, r):
        """Matched filter of chirp_method (OFDM.py:357-358): r [B, T] -> P [B, T+Lc-1], pmax [B]."""
        self._f32(r)
        assert r.dim() == 2
        B, T = r.shape
        rs = r.stride(0) if B > 1 else T
        plen = T + self.chirp_len - 1
        pstride = (plen + 3) // 4 * 4
        P = torch.empty((B, pstride), dtype=torch.float32, device=self.device)
        pmax = torch.empty((B,), dtype=torch.float32, device=self.device)
        wb = int(self.lib.gf3_xcorr_work_bytes(self._plan, B, T))
        work = torch.empty((wb,), dtype=torch.uint8, device=self.device)
        check(self._call("gf3_xcorr", self._plan, _ptr(r), rs, B, T, _ptr(P), pstride, _ptr(pmax), _ptr(work), _STREAM))
        return P[:, :plen], pmax

    def peak_pick(self, P, pmax, T, max_peaks=64):
        """Detection rule of chirp_method (OFDM.py:359-372): -> peaks int64 [B, max_peaks], count int32 [B]."""
        assert P.is_cuda and P.dtype == torch.float32 and P.stride(1) == 1
        B = P.shape[0]
        peaks = torch.full((B, max_peaks), -1, dtype=torch.int64, device=self.device)
        count = torch.empty((B,), dtype=torch.int32, device=self.device)
        work = torch.empty((max(1, int(self.lib.gf3_peak_pick_work_bytes(self._plan, B, T))),), dtype=torch.uint8, device=self.device)
        check(self._call("gf3_peak_pick", self._plan, _ptr(P), P.stride(0), B, T, _ptr(pmax), _ptr(peaks), max_peaks,
                                     _ptr(count), _ptr(work), _STREAM))
        return peaks, count

    _FMT = {torch.uint8: 0, torch.int16: 1, torch.float32: 2}       # GF3_SAMPLE_*

    def sync_streams(self, r, max_peaks=8, detect_only=False):
        """chirp_method (OFDM.py:356-372) for a batch of streams in one call (gf3_sync_streams): r [B, T] float32,
        int16 or uint8 -> (P [B, T+Lc-1], pmax [B], peaks int64 [B, max_peaks], count int32 [B]).
        detect_only (gf3_sync_detect): the same peaks / count / pmax, but P is scratch -- blocks that provably hold no
        candidate are not computed."""
        assert r.is_cuda and r.dim() == 2 and r.stride(1) == 1 and r.dtype in self._FMT
        B, T = r.shape
        rs = r.stride(0) if B > 1 else T
        plen = T + self.chirp_len - 1
        pstride = (plen + 3) // 4 * 4
        P = torch.empty((B, pstride), dtype=torch.float32, device=self.device)
        pmax = torch.empty((B,), dtype=torch.float32, device=self.device)
        peaks = torch.full((B, max_peaks), -1, dtype=torch.int64, device=self.device)
        count = torch.empty((B,), dtype=torch.int32, device=self.device)
        work = torch.empty((max(16, int(self.lib.gf3_sync_work_bytes(self._plan, B, T))),), dtype=torch.uint8, device=self.device)
        check(self._call("gf3_sync_detect" if detect_only else "gf3_sync_streams", self._plan, _ptr(r), self._FMT[r.dtype], rs, B, T, _ptr(P), pstride, _ptr(pmax),
                         _ptr(peaks), max_peaks, _ptr(count), _ptr(work), _STREAM))
        self._last_sync_work = work            # (diagnostics: the first B * ceil(plen / 2048) floats are the block maxima)
        return P[:, :plen], pmax, peaks, count

    def schmidlcox(self, r, search):
        """Schmidl & Cox timing metric of OFDM.py:376-387 for a batch of streams r [B, T]: -> (index int64 [B] = first
        argmax |P| over the first `search` samples, value float64 [B]); the reference returns index + N - 1."""
        assert r.is_cuda and r.dim() == 2 and r.stride(1) == 1 and r.dtype in self._FMT
        B, T = r.shape
        idx = torch.empty((B,), dtype=torch.int64, device=self.device)
        val = torch.empty((B,), dtype=torch.float64, device=self.device)
        check(self._call("gf3_schmidlcox", self._plan, _ptr(r), self._FMT[r.dtype], r.stride(0) if B > 1 else T, B, T, int(search),
                         _ptr(idx), _ptr(val), _STREAM))
        return idx, val

    def peaks_to_offsets(self, peaks, count, r_stride, T, pk_expected):
        """get_symbols' bookkeeping on the device (OFDM.py:393-397): detections of a batch of streams ->
        (pkt_offset int64 [B * pk_expected] into the flat sample array, ok uint8 [B])."""
        B, max_peaks = peaks.shape
        assert peaks.is_cuda and peaks.dtype == torch.int64 and peaks.is_contiguous() and count.dtype == torch.int32
        off = torch.empty((B * pk_expected,), dtype=torch.int64, device=self.device)
        ok = torch.empty((B,), dtype=torch.uint8, device=self.device)
        check(self._call("gf3_peaks_to_offsets", self._plan, _ptr(peaks), _ptr(count), B, max_peaks, r_stride, T, pk_expected,
                         _ptr(off), _ptr(ok), _STREAM))
        return off, ok

    def receive_streams(self, r, pk_expected, xor=True, want_eq=False, out=None):
        """receiver.receive (OFDM.py:581-609) for a batch of raw received streams r float32 [B, T] that hold
        pk_expected packets each: matched filter -> detection rule -> packet offsets -> fused receive chain,
        everything on the device.  -> dict(bits [B * pk_expected, bits_stride], ok uint8 [B] (1: the stream's
        chirps were found where the reference's slicing would succeed), peaks, count, Hs, He, slope[, eq])."""
        assert r.is_cuda and r.dim() == 2 and r.stride(1) == 1 and r.dtype in self._FMT
        B, T = r.shape
        _, _, peaks, count = self.sync_streams(r, pk_expected + 3, detect_only=True)
        off, ok = self.peaks_to_offsets(peaks, count, r.stride(0), T, pk_expected)
        # offsets are relative to r's first sample.  PCM samples (and, when staged is set, float32 ones: packets of a
        # raw stream start at arbitrary sample offsets) enter the receive kernels through the staging buffer
        if r.dtype != torch.float32 or self.staged_streams:
            res, Hs, He, slope = self.rx_receive_pcm(r, B * pk_expected, off, xor=xor, want_eq=want_eq, out=out)
        else:
            res, Hs, He, slope = self.rx_receive(r, B * pk_expected, off, xor=xor, want_eq=want_eq, out=out)
        d = dict(bits=res[0] if want_eq else res, ok=ok, peaks=peaks, count=count, Hs=Hs, He=He, slope=slope, pkt_offset=off)
        if want_eq:
            d["eq"] = res[1]
        return d

    # ------------------------------------------------------------------ transmit chain
    def tx_len(self, pk_per_stream):
        return pk_per_stream * (self.chirp_len + self.pkt_samples) + self.chirp_len

    def tx_modulate(self, bits_packed, filler, n_streams, pk_per_stream, out=None, xor=False):
        """bits_packed uint8 [n_streams, pk_per_stream, bits_stride] (encoded bits, MSB first; with xor=True the
        UN-encoded bits: encode("XOR") is fused into the kernel), filler complex64 [n_streams, K-Nd]
        -> waveform float32 [n_streams, tx_len]  (OFDM.py:163-166, 191-226, 244-259, 322-323)."""
        assert bits_packed.is_cuda and bits_packed.dtype == torch.uint8 and bits_packed.is_contiguous()
        stride = bits_packed.shape[-1]
        if filler is not None:
            assert filler.is_cuda and filler.dtype == torch.complex64 and filler.is_contiguous()
        T = self.tx_len(pk_per_stream)
        tstride = (T + 3) // 4 * 4
        if out is None:
            out = torch.empty((n_streams, tstride), dtype=torch.float32, device=self.device)
        if xor:
            check(self._call("gf3_tx_encode_modulate", self._plan, _ptr(bits_packed), stride, _ptr(self.xor2), _ptr(filler), _ptr(self.known),
                             n_streams, pk_per_stream, _ptr(out), out.stride(0), _STREAM))
        else:
            check(self._call("gf3_tx_modulate", self._plan, _ptr(bits_packed), stride, _ptr(filler), _ptr(self.known),
                             n_streams, pk_per_stream, _ptr(out), out.stride(0), _STREAM))
        return out[:, :T]

    def ifft_symbols(self, spectrum):
        """np.fft.ifft of Hermitian spectra given as bins 1..K: complex64 [n, K] -> float32 [n, N + cp] (CP prepended, no gain)."""
        assert spectrum.is_cuda and spectrum.dtype == torch.complex64 and spectrum.is_contiguous() and spectrum.shape[-1] == self.K
        n = spectrum.numel() // self.K
        out = torch.empty((n, self.symlen), dtype=torch.float32, device=self.device)
        check(self._call("gf3_tx_ifft", self._plan, _ptr(spectrum), n, _ptr(out), _STREAM))
        return out

    def cdiv(self, Y, H):
        """Y / H, H broadcast over rows: complex64 [n, m] / [m]."""
        assert Y.is_cuda and Y.dtype == torch.complex64 and Y.is_contiguous() and H.is_cuda and H.dtype == torch.complex64
        m = Y.shape[-1]
        assert H.numel() == m
        out = torch.empty_like(Y)
        check(self._call("gf3_cdiv", _ptr(Y), _ptr(H.contiguous()), Y.numel() // m, m, _ptr(out), _STREAM))
        return out

    # ------------------------------------------------------------------ channel + counters
    def channel_sim(self, x, taps, sigma, seed, stream_ids=None):
        """y = lfilter(taps, 1, x) + sigma*N(0,1): x [B, T] (row stride may exceed T).  stream_ids (int64 [B]):
        the noise of a row then depends on (seed, id, sample) only, not on the row's place in the batch."""
        assert x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
        B, T = x.shape
        taps = taps.to(torch.float32).contiguous()
        y = torch.empty((B, (T + 3) // 4 * 4), dtype=torch.float32, device=self.device)
        sg = sigma.to(torch.float32).contiguous() if sigma is not None else None
        if stream_ids is None:
            check(self._call("gf3_channel_sim", _ptr(x), x.stride(0), B, T, _ptr(taps), taps.shape[1], _ptr(sg),
                             int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(y), y.stride(0), _STREAM))
        else:
            ids = stream_ids.to(device=self.device, dtype=torch.int64).contiguous()
            assert ids.numel() == B
            check(self._call("gf3_channel_sim_ids", _ptr(x), x.stride(0), B, T, _ptr(taps), taps.shape[1], _ptr(sg), _ptr(ids),
                             int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(y), y.stride(0), _STREAM))
        return y[:, :T]

    def random_bytes(self, n_rows, row_bytes, seed, row_ids=None, out=None):
        """Uniform random bytes uint8 [n_rows, row_bytes] (Philox; row r from the stream of row_ids[r])."""
        if out is None:
            out = torch.empty((n_rows, row_bytes), dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.stride(1) == 1 and out.shape == (n_rows, row_bytes)
        ids = row_ids.to(device=self.device, dtype=torch.int64).contiguous() if row_ids is not None else None
        for r0 in range(0, n_rows, 65535):
            n = min(65535, n_rows - r0)
            check(self._call("gf3_random_bytes", _ptr(out[r0:]), out.stride(0), n, row_bytes,
                             _ptr(ids[r0:]) if ids is not None else None, int(seed) & 0xFFFFFFFFFFFFFFFF, _STREAM))
        return out

    def ber_count(self, a, b, nbits, counter):
        """counter (int64/uint64 [2]) += (bit errors, bits) between two packed rows."""
        check(self._call("gf3_ber_count", _ptr(a), _ptr(b), nbits, _ptr(counter), _STREAM))
        return counter

    def pcm_to_f32(self, pcm, out=None):
        """uint8 / int16 PCM device tensor -> float32 (exact values, no DC removal)."""
        assert pcm.is_cuda and pcm.is_contiguous() and pcm.dtype in (torch.uint8, torch.int16)
        if out is None:
            out = torch.empty(pcm.shape, dtype=torch.float32, device=self.device)
        check(self._call("gf3_pcm_to_f32", _ptr(pcm), 0 if pcm.dtype == torch.uint8 else 1, pcm.numel(), _ptr(out), _STREAM))
        return out

    # ------------------------------------------------------------------ stage-level methods
    def eq_estimate(self, start, end):
        """First half of receiver.equalise on spectra (OFDM.py:429-462):
        start, end complex64 [n_packets, P, K] -> Hs, He [n_packets, K], slope float64 [n_packets]."""
        assert start.is_cuda and start.dtype == torch.complex64 and start.is_contiguous() and start.shape[1:] == (self.P, self.K)
        assert end.is_cuda and end.dtype == torch.complex64 and end.is_contiguous() and end.shape == start.shape
        n = start.shape[0]
        Hs = torch.empty((n, self.K), dtype=torch.complex64, device=self.device)
        He = torch.empty_like(Hs)
        slope = torch.empty((n,), dtype=torch.float64, device=self.device)
        check(self._call("gf3_eq_estimate", self._plan, _ptr(start), _ptr(end), n, _ptr(self.known), _ptr(Hs), _ptr(He),
                                       _ptr(slope), _STREAM))
        return Hs, He, slope

    def eq_apply(self, data, Hs, He, slope, want_hest=True):
        """Second half of receiver.equalise (OFDM.py:466-478): data complex64 [n_packets, L, K]
        -> eq (and Hest) complex64 [n_packets, L, K]."""
        assert data.is_cuda and data.dtype == torch.complex64 and data.is_contiguous() and data.shape[1:] == (self.L, self.K)
        n = data.shape[0]
        eq = torch.empty_like(data)
        hest = torch.empty_like(data) if want_hest else None
        check(self._call("gf3_eq_apply", self._plan, _ptr(data), n, _ptr(Hs.contiguous()), _ptr(He.contiguous()),
                                    _ptr(slope.contiguous()), _ptr(eq), _ptr(hest), _STREAM))
        return (eq, hest) if want_hest else eq

    def demap(self, symbols, want_hard=True):
        """receiver.demap (OFDM.py:484-500): complex64 [...] -> bits uint8 [..., 2] (and hard decisions)."""
        assert symbols.is_cuda and symbols.dtype == torch.complex64 and symbols.is_contiguous()
        n = symbols.numel()
        bits = torch.empty(symbols.shape + (2,), dtype=torch.uint8, device=self.device)
        hard = torch.empty_like(symbols) if want_hard else None
        check(self._call("gf3_demap", _ptr(symbols), n, _ptr(bits), _ptr(hard), _STREAM))
        return (bits, hard) if want_hard else bits

    def tx_frame(self, data_time, sync):
        """transmitter.send_to_stream (OFDM.py:244-259): data_time float32 [n_packets, L*(N+cp)],
        sync float32 [Ls] -> framed waveform float32 [n_packets*(Ls + (2P+L)(N+cp)) + Ls]."""
        assert data_time.is_cuda and data_time.dtype == torch.float32 and data_time.is_contiguous()
        assert sync.is_cuda and sync.dtype == torch.float32 and sync.is_contiguous()
        n = data_time.shape[0]
        ls = sync.numel()
        out = torch.empty((n * (ls + self.pkt_samples) + ls,), dtype=torch.float32, device=self.device)
        check(self._call("gf3_tx_frame", self._plan, _ptr(data_time), n, _ptr(sync), ls, _ptr(self.known), _ptr(out), _STREAM))
        return out

    # ------------------------------------------------------------------ whole receive chain
    def rx_receive(self, samples, n_packets, pkt_offset=None, xor=True, want_eq=False, out=None):
        """Channel estimate + data symbols of n_packets packets in ONE launch (gf3_rx_receive):
        -> (packed bits [n_packets, bits_stride] (, eq), Hs, He, slope)."""
        self._f32(samples)
        off = self._offsets(pkt_offset, n_packets)
        Hs = torch.empty((n_packets, self.K), dtype=torch.complex64, device=self.device)
        He = torch.empty_like(Hs)
        slope = torch.empty((n_packets,), dtype=torch.float64, device=self.device)
        bits = out if out is not None else torch.empty((n_packets, self.bits_stride), dtype=torch.uint8, device=self.device)
        eq = torch.empty((n_packets, self.L, self.K), dtype=torch.complex64, device=self.device) if want_eq else None
        check(self._call("gf3_rx_receive", self._plan, _ptr(samples), _ptr(off), n_packets, _ptr(self.known), _ptr(Hs), _ptr(He),
                                      _ptr(slope), _ptr(self.xor2) if xor else None, _ptr(bits), self.bits_stride, _ptr(eq),
                                      _STREAM))
        return ((bits, eq) if want_eq else bits), Hs, He, slope

    def rx_receive_pcm(self, samples, n_packets, pkt_offset=None, xor=True, want_eq=False, out=None):
        """rx_receive on samples in their recorded format (uint8 / int16 PCM, or float32): the symbols enter the
        kernels through the cp.async.bulk staging buffer and are converted in registers (gf3_rx_receive_pcm).
        -> (packed bits [n_packets, bits_stride] (, eq), Hs, He, slope)."""
        assert samples.is_cuda and samples.stride(-1) == 1 and samples.dtype in self._FMT
        off = self._offsets(pkt_offset, n_packets)
        Hs = torch.empty((n_packets, self.K), dtype=torch.complex64, device=self.device)
        He = torch.empty_like(Hs)
        slope = torch.empty((n_packets,), dtype=torch.float64, device=self.device)
        bits = out if out is not None else torch.empty((n_packets, self.bits_stride), dtype=torch.uint8, device=self.device)
        eq = torch.empty((n_packets, self.L, self.K), dtype=torch.complex64, device=self.device) if want_eq else None
        check(self._call("gf3_rx_receive_pcm", self._plan, _ptr(samples), self._FMT[samples.dtype], _ptr(off), n_packets,
                         _ptr(self.known), _ptr(Hs), _ptr(He), _ptr(slope), _ptr(self.xor2) if xor else None, _ptr(bits),
                         self.bits_stride, _ptr(eq), _STREAM))
        return ((bits, eq) if want_eq else bits), Hs, He, slope
