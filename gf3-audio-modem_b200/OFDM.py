"""Drop-in replacement for the reference's OFDM.py (adamg-97/GF3-audio-modem).

Same classes, methods, argument meaning, prints and error behaviour as the reference module, so
`from OFDM import *` in the notebooks keeps working (Final System Test.ipynb:9) -- but the
physical-layer arithmetic (transmit chain, chirp synchronisation, receive chain) runs in
hand-written sm_100a CUDA kernels (libgf3b200.so) through gf3b200.Phy.  numpy arrays in, bits or
waveforms out.  No CPU fallback: without a CUDA device the compute methods raise gf3b200.Gf3Error.

Host-side pieces kept in numpy are index / RNG bookkeeping only: bit padding and XOR coding
(np.random draw order must match the reference, OFDM.py:172,203), reshapes, slicing, file framing.
"""
import os
import sys

import numpy as np
import scipy  # noqa: F401  (re-exported like the reference, OFDM.py:3)
from scipy.io import wavfile  # noqa: F401
from scipy.signal import chirp, convolve, lfilter  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import gf3b200  # noqa: E402
from gf3b200 import Phy  # noqa: E402


class _Missing:
    """Stand-in for an optional module of the reference (matplotlib, sounddevice, IPython, pyldpc)
    that is not installed: attribute access raises with a clear message."""

    def __init__(self, name):
        object.__setattr__(self, "_name", name)

    def __getattr__(self, item):
        raise ImportError("optional dependency '%s' is not installed (needed for %s.%s)" % (self._name, self._name, item))

    def __setattr__(self, k, v):
        pass


def _optional(name, attr=None):
    try:
        mod = __import__(name, fromlist=["_"])
        return getattr(mod, attr) if attr else mod
    except Exception:
        return _Missing(name)


plt = _optional("matplotlib.pyplot")          # OFDM.py:4
sd = _optional("sounddevice")                 # OFDM.py:7
try:
    sd.default.channels = 1                   # OFDM.py:8
except Exception:
    pass
Audio = _optional("IPython.display", "Audio")  # OFDM.py:9
pyldpc = _optional("pyldpc")                  # OFDM.py:10
pd = _optional("pandas")                      # re-exported by the revisions the older notebooks ran against (Weekend Challenge.ipynb:33)


def _find_ci(path):
    """Resolve a relative path case-insensitively (the reference opens handouts/, input_files/
    but the repository has Handouts/, input_Files/: OFDM.py:99,757)."""
    if os.path.exists(path):
        return path
    cur = "." if not os.path.isabs(path) else os.sep
    for part in [p for p in path.split(os.sep) if p]:
        try:
            entries = os.listdir(cur)
        except OSError:
            return path
        match = [e for e in entries if e.lower() == part.lower()]
        if not match:
            return path
        cur = os.path.join(cur, match[0])
    return cur


#########################################
#   Old API (older OFDM.py revisions)   #
#########################################
# The Weekend-Challenge / Week-2 / Initial-test / Audio notebooks were written against earlier revisions of OFDM.py
# whose code is not in the reference repository: CamG(N, cp, "QPSK") with K = N, module-level FFT / IFFT /
# equalise(Y, H), keyword receiver(ofdm_symbol_size=, cp_length=, ...).  These shims give that surface on top of
# the same device kernels, so the notebooks' cells run unmodified.

_legacy_phys = {}


def _legacy_phy(N, cp=0):
    """Full-band plan without known symbols for an N-point symbol (cached)."""
    N, cp = int(N), int(cp)
    if N < 64 or N > 4096 or N & (N - 1):
        raise ValueError("symbol size %d: the device FFT covers powers of two from 64 to 4096 (there is no CPU path)" % N)
    key = (N, cp)
    if key not in _legacy_phys:
        _legacy_phys[key] = Phy(N=N, cp=cp, lo=1, hi=N // 2, n_pilots=0, packet_len=1)
    return _legacy_phys[key]


def _as_real_rows(x, what):
    x = np.asarray(x)
    if np.iscomplexobj(x):
        if np.max(np.abs(x.imag), initial=0.0) > 1e-9 * max(1.0, np.max(np.abs(x.real), initial=0.0)):
            raise ValueError("%s: the device transform takes real samples" % what)
        x = x.real
    return np.ascontiguousarray(x, dtype=np.float64)


def FFT(x):
    """np.fft.fft of real symbols along the last axis (old API; Weekend Challenge.ipynb:195): bins 1..N/2-1 come from
    the device kernel (gf3_rx_spectrum), the mirror half is their conjugate, DC / Nyquist are two sums."""
    import torch
    x = _as_real_rows(x, "FFT")
    N = x.shape[-1]
    phy = _legacy_phy(N)
    rows = x.reshape(-1, N)
    spec = phy.spectrum(torch.from_numpy(rows.astype(np.float32).reshape(-1)).to(phy.device), rows.shape[0]).cpu().numpy().astype(np.complex128)
    out = np.zeros((rows.shape[0], N), dtype=np.complex128)
    out[:, 1:N // 2] = spec
    out[:, N // 2 + 1:] = np.conj(spec[:, ::-1])
    out[:, 0] = rows.sum(axis=1)
    out[:, N // 2] = rows[:, ::2].sum(axis=1) - rows[:, 1::2].sum(axis=1)
    return out.reshape(x.shape)


def IFFT(X):
    """np.fft.ifft of Hermitian-symmetric spectra along the last axis (old API; Initial OFDM Test.ipynb cell 13): bins
    1..N/2-1 go through the device transmit kernel (gf3_tx_ifft); a DC / Nyquist term is a constant / alternating offset."""
    import torch
    X = np.asarray(X)
    N = X.shape[-1]
    phy = _legacy_phy(N)
    rows = X.reshape(-1, N).astype(np.complex128)
    half = np.ascontiguousarray(rows[:, 1:N // 2].astype(np.complex64))
    if not np.allclose(rows[:, N // 2 + 1:], np.conj(rows[:, 1:N // 2][:, ::-1]), rtol=1e-6, atol=1e-9 * (1.0 + np.max(np.abs(rows)))):
        raise ValueError("IFFT: the device transform takes Hermitian-symmetric spectra (real symbols)")
    t = phy.ifft_symbols(torch.from_numpy(half).to(phy.device)).cpu().numpy().astype(np.float64)
    t += rows[:, :1].real / N + rows[:, N // 2:N // 2 + 1].real / N * np.where(np.arange(N) % 2 == 0, 1.0, -1.0)[None, :]
    return t.reshape(X.shape).astype(np.complex128)


def equalise(Y, H):
    """Old API (Weekend Challenge.ipynb:225): one-tap equaliser with a KNOWN channel, Y / H on the device."""
    import torch
    Y = np.asarray(Y)
    H = np.asarray(H)
    phy = _legacy_phy(64)                      # any plan: the divide needs none of its tables
    rows = np.ascontiguousarray(Y.reshape(-1, Y.shape[-1]).astype(np.complex64))
    out = phy.cdiv(torch.from_numpy(rows).to(phy.device), torch.from_numpy(np.ascontiguousarray(H.reshape(-1).astype(np.complex64))).to(phy.device))
    return out.cpu().numpy().astype(np.complex128).reshape(Y.shape)


def _is_old_ctor(mode):
    return isinstance(mode, (int, np.integer)) and not isinstance(mode, bool)


class _OldCamG:
    """CamG(N, cp, "QPSK") of the old API (Weekend Challenge.ipynb:75, Initial OFDM Test.ipynb:34): K is the FFT
    length, every bin 1..N/2-1 carries data, and the stage helpers live on the parameter object itself."""

    def __init__(self, N, cp, modulation="QPSK"):
        if modulation != "QPSK":
            raise ValueError("Invalid Modulation Type")
        self.ofdm_symbol_size = self.K = int(N)
        self.cp_length = int(cp)
        self.modulation, self.mu, self.fs = modulation, 2, 48000
        self.all_carriers = np.arange(self.K)
        self.data_carriers = np.arange(1, self.K // 2)
        self.bits_per_symbol = len(self.data_carriers) * self.mu
        self.mapping_table = {(0, 0): (1 + 1j) / np.sqrt(2), (1, 0): (1 - 1j) / np.sqrt(2),
                              (1, 1): (-1 - 1j) / np.sqrt(2), (0, 1): (-1 + 1j) / np.sqrt(2)}

    def SP(self, bits):
        return np.asarray(bits).reshape(-1, self.mu)

    def map(self, bits):
        return gf3b200.qpsk_points(bits)

    def OFDM_symbol(self, payload):
        payload = np.asarray(payload).reshape(-1, len(self.data_carriers))
        sym = np.zeros((payload.shape[0], self.K), dtype=complex)
        sym[:, self.data_carriers] = payload
        sym[:, -self.data_carriers] = np.conj(payload)
        return sym[0] if sym.shape[0] == 1 else sym

    def add_cp(self, time_data):
        time_data = np.asarray(time_data)
        if self.cp_length == 0:
            return time_data
        return np.concatenate([time_data[..., -self.cp_length:], time_data], axis=-1)

    def remove_cp(self, rx):
        return np.asarray(rx)[..., self.cp_length:]

    def get_data(self, spectrum):
        return np.asarray(spectrum)[..., 1:self.K // 2]

    def demap(self, symbols):
        """Minimum-distance QPSK decisions on the device (gf3_demap) -> (bits[..., 2], hardDecision)."""
        import torch
        if type(symbols) != np.ndarray:                            # noqa: E721
            raise ValueError("Symbols must be numpy array")
        phy = _legacy_phy(max(64, self.K))
        bits, hard = phy.demap(torch.from_numpy(np.ascontiguousarray(symbols, dtype=np.complex64)).to(phy.device))
        return bits.cpu().numpy().astype(np.int64), hard.cpu().numpy().astype(np.complex128)

    def PS(self, bits):
        return np.asarray(bits).reshape((-1,))

    def __repr__(self):
        return "Number of Sub Carriers: {} \nCyclic prefix length: {} \nModulation method: {}".format(self.K, self.cp_length, self.modulation)


#########################################
#                 CamG                  #
#########################################

class CamG:
    """Parameter object (OFDM.py:17-115).  Extra keyword arguments (not in the reference) reach
    the parameter space its notebooks used through older revisions: any power-of-two symbol size,
    any CP, any data-bin range."""

    def __new__(cls, mode=None, *args, **kwargs):
        if cls is CamG and _is_old_ctor(mode):                     # CamG(N, cp, "QPSK"): the old parameter object
            return _OldCamG(mode, *args, **kwargs)
        return super().__new__(cls)

    def __init__(self, mode=None, encoding="None", no_pilots=20, packet_length=180, *,
                 ofdm_symbol_size=None, cp_length=None, lowest_bin=None, highest_bin=None,
                 modulation=None, fs=None, end_sync=True, pilot_sequence=None, sync_method="chirp"):
        # old API of transmitter / receiver: (N, cp, "QPSK") positionally (Audio.ipynb:60,147) or ofdm_symbol_size= / cp_length= /
        # modulation= / fs= / end_sync= / no_pilots= / pilot_sequence= / sync_method= by keyword (Week 2 Challenge.ipynb:42-45,126,305):
        # every bin 1..N/2-1 carries data, no bit coding
        self.old_api = _is_old_ctor(mode) or (mode is None and ofdm_symbol_size is not None)
        if _is_old_ctor(mode):
            ofdm_symbol_size, cp_length = int(mode), int(encoding)
            if isinstance(no_pilots, str):
                modulation, no_pilots = no_pilots, 20
            mode, encoding = None, "None"
        if self.old_api:
            mode = "A1"
            lowest_bin = 1 if lowest_bin is None else lowest_bin
            highest_bin = int(ofdm_symbol_size) // 2 if highest_bin is None else highest_bin
        if mode is None:
            raise TypeError("CamG() missing required argument: 'mode'")
        if modulation is not None and modulation != "QPSK":
            raise ValueError("Invalid Modulation Type")
        self.end_sync, self.gap_length, self.Hest = end_sync, 0, None      # old-API attributes (Week 2 Challenge.ipynb:45,305)
        self.encoding = encoding
        self.fs = 48000 if fs is None else fs
        self.ofdm_symbol_size = 4096 if ofdm_symbol_size is None else int(ofdm_symbol_size)
        self.K = self.ofdm_symbol_size // 2 - 1
        K = 2047
        modes = {
            "A1": (224, (1, K)), "A2": (224, (100, 1500)), "A3": (224, (100, 1000)),
            "B1": (704, (1, K)), "B2": (704, (100, 1500)), "B3": (704, (100, 1000)),
            "C1": (1184, (1, K)), "C2": (1184, (100, 1500)), "C3": (1184, (100, 1000)),
        }
        self.cp_length = modes[mode][0] if cp_length is None else int(cp_length)
        self.lowest_bin = modes[mode][1][0] if lowest_bin is None else int(lowest_bin)
        self.highest_bin = modes[mode][1][1] if highest_bin is None else int(highest_bin)
        self.packet_length = packet_length
        self.no_pilots = no_pilots
        self.sync_method = sync_method
        self.L = self.K + 1
        self.f0 = 0
        self.f1 = 8000
        self.modulation = "QPSK"
        if self.modulation == "QPSK":
            self.mapping_table = {
                (0, 0): (1 + 1j) / np.sqrt(2), (1, 0): (1 - 1j) / np.sqrt(2),
                (1, 1): (-1 - 1j) / np.sqrt(2), (0, 1): (-1 + 1j) / np.sqrt(2),
            }
            self.mu = 2
        else:
            raise ValueError("Invalid Modulation Type")
        self._derive()
        # known sequence: handouts/random_bits.txt if the caller's tree has it, else the packaged copy
        path = _find_ci(os.path.join("handouts", "random_bits.txt"))
        if os.path.exists(path):
            with open(path, "rb") as f:
                raw = np.frombuffer(f.read(4096), dtype=np.uint8)
            self.known_sequence = (raw - ord("0")).astype(np.int64)
        else:
            self.known_sequence = gf3b200.default_known_sequence()
        if pilot_sequence is not None:                             # old API: the caller's known bits (Audio.ipynb:147)
            self.known_sequence = np.asarray(pilot_sequence).astype(np.int64).reshape(-1)
        self.pilot_sequence = self.known_sequence
        self._phy = None
        self._phy_key = None

    def _derive(self):
        """Derived attributes of OFDM.py:46-49,64,94-95 (recomputed when the base ones change)."""
        self.carriers = np.arange(1, self.K + 1)
        self.data_carriers = np.arange(self.lowest_bin, self.highest_bin)
        self.data_carriers_per_symbol = len(self.data_carriers)
        self.unused_carriers = np.delete(self.carriers, (self.data_carriers - 1))
        self.chirp_length = 5 * (self.ofdm_symbol_size + self.cp_length)
        self.data_bits_per_symbol = self.data_carriers_per_symbol * self.mu
        self.bits_per_symbol = self.K * self.mu

    @property
    def phy(self):
        """The device plan for the CURRENT attribute values (rebuilt if they were patched)."""
        if self.pilot_sequence is not self.known_sequence and self.old_api:     # old API: `tx.pilot_sequence = bits` after construction
            self.known_sequence = np.asarray(self.pilot_sequence).astype(np.int64).reshape(-1)
            self.pilot_sequence = self.known_sequence
        key = (self.ofdm_symbol_size, self.cp_length, self.lowest_bin, self.highest_bin, self.no_pilots,
               self.packet_length, self.chirp_length, self.f0, self.f1, self.fs, self.known_sequence.tobytes()[:64], len(self.known_sequence))
        if self._phy is None or key != self._phy_key:
            self._phy = Phy(N=self.ofdm_symbol_size, cp=self.cp_length, lo=self.lowest_bin, hi=self.highest_bin,
                            n_pilots=self.no_pilots, packet_len=self.packet_length,
                            known_sequence=self.known_sequence, chirp_len=self.chirp_length,
                            fs=float(self.fs), f0=float(self.f0), f1=float(self.f1))
            self._phy_key = key
        return self._phy

    def sync_chirp(self):
        """OFDM.py:106-109 (device kernel, float64 out)."""
        return self.phy.sync_chirp().cpu().numpy().astype(np.float64)

    def __repr__(self):
        return ("Number of actual Sub Carriers:      {:.0f} \nCyclic prefix length:               {:.0f} \n"
                "Modulation method:                  {} \nSync Method:                        {} \n"
                "Packet Length:                      {}").format(self.K, self.cp_length, self.modulation,
                                                                 self.sync_method, self.packet_length)


#########################################
#              Transmitter              #
#########################################

class transmitter(CamG):

    def encode(self, bits):
        """OFDM.py:128-187.  Host side: the RNG draw (np.random.binomial) must stay in the
        reference's order so seeded runs see identical padding."""
        if self.encoding == "LDPC":
            raise NotImplementedError('encoding "LDPC" is marked broken in the reference (OFDM.py:21) and is out of scope')
        bits = np.asarray(bits)
        if self.encoding == "XOR":
            dbs = self.data_bits_per_symbol
            known_bits = np.tile(self.known_sequence[:dbs], int(np.ceil(len(bits) / dbs)))[:len(bits)]
            bits = np.bitwise_xor(bits, known_bits)
        bits_per_packet = self.data_bits_per_symbol * self.packet_length
        padding_length = (bits_per_packet - len(bits) % bits_per_packet) % bits_per_packet
        padding = np.random.binomial(n=1, p=0.5, size=(padding_length,))
        return np.hstack([bits, padding])

    def SP(self, bits):
        return bits.reshape(-1, self.data_carriers_per_symbol, self.mu)

    def map(self, bits):
        """OFDM.py:196-197 as a table lookup (formatting only; transmit() maps on the device)."""
        return gf3b200.qpsk_points(bits)

    def random_qpsk(self):
        qpsk = np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
        return np.random.choice(qpsk, size=(self.K - self.data_carriers_per_symbol), replace=True)

    def build_OFDM_symbol(self, payload):
        symbols = np.zeros([payload.shape[0], self.ofdm_symbol_size], dtype=complex)
        rand_qpsk = self.random_qpsk()
        symbols[:, self.data_carriers] = payload
        symbols[:, self.unused_carriers] = rand_qpsk
        symbols[:, -self.data_carriers] = np.conj(payload)
        symbols[:, -self.unused_carriers] = np.conj(rand_qpsk)
        return symbols

    def add_cp(self, time_data):
        if self.cp_length == 0:
            return time_data
        return np.hstack([time_data[:, -self.cp_length:], time_data])

    def build_schmidlcox(self):
        """OFDM.py:230-238 (host-side formatting; like the reference it only works for geometries whose known sequence
        reshapes into whole symbols)."""
        symbols = self.map(self.SP(self.known_sequence))
        p = np.zeros(self.K, dtype=complex)
        p[::2] = symbols[0, :self.K // 2]
        return p.reshape(-1, self.K)

    def send_to_stream(self, time_data, sync):
        """OFDM.py:242-276: frame time-domain symbols (CP included) into packets with the known
        symbols and the caller's sync waveform; the frame is assembled on the device."""
        import torch
        phy = self.phy
        symlen = self.ofdm_symbol_size + self.cp_length
        packets = np.asarray(time_data).reshape(-1, self.packet_length, symlen)             # OFDM.py:251 (raises like numpy)
        self.no_packets = packets.shape[0]
        sync = np.asarray(sync)
        d_data = torch.from_numpy(np.ascontiguousarray(packets.real, dtype=np.float32).reshape(self.no_packets, -1)).to(phy.device)
        d_sync = torch.from_numpy(np.ascontiguousarray(sync.real, dtype=np.float32)).to(phy.device)
        tx = phy.tx_frame(d_data, d_sync).cpu().numpy().astype(np.float64)
        # frames for visuals (OFDM.py:262-274): index bookkeeping on the host
        ls = sync.shape[0]
        frame_length = tx.shape[0]
        sync_valid = np.zeros(frame_length); known_valid = np.zeros(frame_length); payload_valid = np.zeros(frame_length)
        sync_valid[0:ls] = 1
        if ls:
            sync_valid[-ls:] = 1
        for f in np.hstack([np.arange(self.no_pilots), np.arange(self.no_pilots + self.packet_length, 2 * self.no_pilots + self.packet_length)]):
            known_valid[ls + symlen * f + self.cp_length + np.arange(self.ofdm_symbol_size)] = 1
        for f in range(self.packet_length):
            payload_valid[ls + symlen * (self.no_pilots + f) + self.cp_length + np.arange(self.ofdm_symbol_size)] = 1
        sync_valid = np.tile(sync_valid, self.no_packets); known_valid = np.tile(known_valid, self.no_packets)
        payload_valid = np.tile(payload_valid, self.no_packets)
        return tx, sync_valid, known_valid, payload_valid

    def _modulate(self, bits_encoded, filler, device_xor=False):
        """Fused device transmit chain: (encoded) bits -> framed waveform (float64 numpy); device_xor: the bits are
        un-encoded and encode("XOR") runs inside the kernel."""
        import torch
        phy = self.phy
        bpp = phy.bits_per_packet
        if len(bits_encoded) % bpp:
            raise ValueError("cannot reshape array of size %d into packets of %d bits" % (len(bits_encoded), bpp))
        n_packets = len(bits_encoded) // bpp
        self.no_packets = n_packets
        packed = np.zeros((n_packets, phy.bits_stride), dtype=np.uint8)
        pk = np.packbits(np.asarray(bits_encoded, dtype=np.uint8).reshape(n_packets, bpp), axis=1)
        packed[:, : pk.shape[1]] = pk
        d_bits = torch.from_numpy(packed).to(phy.device).reshape(1, n_packets, phy.bits_stride)
        d_fill = torch.from_numpy(np.asarray(filler).astype(np.complex64)).to(phy.device).reshape(1, -1) if phy.K > phy.Nd else None
        out = phy.tx_modulate(d_bits, d_fill, 1, n_packets, xor=device_xor)
        return out[0].cpu().numpy().astype(np.float64)

    def graphs(self):
        """OFDM.py:279-292: plot of the Gray-mapped QPSK constellation (needs matplotlib)."""
        for b1 in [0, 1]:
            for b0 in [0, 1]:
                B = (b1, b0)
                Q = self.mapping_table[B]
                plt.plot(Q.real, Q.imag, 'bo')
                plt.text(Q.real, Q.imag + 0.1, "".join(str(x) for x in B), ha='center')
        plt.grid(alpha=0.5)
        plt.xlim(-1, 1)
        plt.ylim(-1, 1)
        plt.title("QPSK Constellation with Gray Mapping")
        plt.show()

    def _pad_for_device_encode(self, bits):
        """encode() with the XOR left to the device (OFDM.py:163-173): the padding is drawn exactly as encode() draws it
        (same np.random.binomial call) and pre-XORed with the known bits, so that the kernel's XOR of the WHOLE packet
        (gf3_tx_encode_modulate) leaves the reference's un-encoded padding behind it: x ^ k ^ k = x."""
        bits = np.asarray(bits)
        bits_per_packet = self.data_bits_per_symbol * self.packet_length
        padding_length = (bits_per_packet - len(bits) % bits_per_packet) % bits_per_packet
        padding = np.random.binomial(n=1, p=0.5, size=(padding_length,))
        dbs = self.data_bits_per_symbol
        pos = (len(bits) + np.arange(padding_length)) % dbs
        return np.hstack([bits, np.bitwise_xor(padding, self.known_sequence[:dbs][pos])])

    def _modulate_packed(self, rows, filler, device_xor):
        """rows uint8 [n_packets, bits_per_packet / 8] (MSB-first bytes, np.packbits order) -> framed waveform."""
        import torch
        phy = self.phy
        n_packets = rows.shape[0]
        self.no_packets = n_packets
        d_bits = torch.zeros((1, n_packets, phy.bits_stride), dtype=torch.uint8, device=phy.device)
        d_bits[0, :, : rows.shape[1]].copy_(torch.from_numpy(rows))
        d_fill = torch.from_numpy(np.asarray(filler).astype(np.complex64)).to(phy.device).reshape(1, -1) if phy.K > phy.Nd else None
        out = phy.tx_modulate(d_bits, d_fill, 1, n_packets, xor=device_xor)
        return out[0].cpu().numpy().astype(np.float64)

    def transmit_bytes(self, payload, graph_output=False):
        """transmit(np.unpackbits(payload)) without the host bit array (file framing + XOR coding on the
        device): the file's bytes ARE the packed bit stream (np.packbits order, OFDM.py:761), the XOR encode runs in the
        transmit kernel, only the padding (< one packet, drawn with the reference's np.random.binomial call so seeded runs
        agree with transmit()) is built from bits on the host.  Same waveform as transmit() (tested)."""
        payload = np.ascontiguousarray(np.asarray(payload, dtype=np.uint8).reshape(-1))
        bpp = self.data_bits_per_symbol * self.packet_length
        if bpp % 8 or self.encoding not in ("XOR", "None"):          # packets not byte aligned: the bit path handles it
            return self.transmit(np.unpackbits(payload), graph_output)
        print("-" * 42 + "\nTRANSMIT\n" + "-" * 42)
        print("OFDM Paramters:")
        print(self)
        nbits = 8 * len(payload)
        padding_length = (bpp - nbits % bpp) % bpp
        padding = np.random.binomial(n=1, p=0.5, size=(padding_length,))          # OFDM.py:172 / :183, same draw
        device_xor = self.encoding == "XOR"
        if device_xor:                                                             # see _pad_for_device_encode
            dbs = self.data_bits_per_symbol
            padding = np.bitwise_xor(padding, self.known_sequence[:dbs][(nbits + np.arange(padding_length)) % dbs])
        filler = self.random_qpsk()
        rows = np.concatenate([payload, np.packbits(padding.astype(np.uint8))]).reshape(-1, bpp // 8)
        print("Number of bits to transmit:         " + str(nbits))
        print("Number of OFDM symbols to transmit: " + str(rows.shape[0] * self.packet_length))
        signal = self._modulate_packed(rows, filler, device_xor)
        print("Number of packets to transmit:      " + str(self.no_packets))
        return signal

    def transmit(self, bits, graph_output=False):
        """OFDM.py:296-343."""
        print("-" * 42 + "\nTRANSMIT\n" + "-" * 42)
        print("OFDM Paramters:")
        print(self)
        if self.sync_method != "chirp":
            raise NotImplementedError("sync_method %r: only the chirp synchronisation is built (OFDM.py:56 hard-wires it; the reference "
                                      "marks schmidlcox_method as no longer working, OFDM.py:376)" % (self.sync_method,))
        device_xor = self.encoding == "XOR"
        bits_encoded = self._pad_for_device_encode(bits) if device_xor else self.encode(bits)
        filler = self.random_qpsk()                     # same draw order as build_OFDM_symbol (OFDM.py:210)
        n_symbols = len(bits_encoded) // self.data_bits_per_symbol
        print("Number of bits to transmit:         " + str(len(bits)))
        print("Number of OFDM symbols to transmit: " + str(n_symbols))
        signal = self._modulate(bits_encoded, filler, device_xor)
        print("Number of packets to transmit:      " + str(self.no_packets))
        if graph_output:
            time = np.linspace(0, len(signal) / self.fs, len(signal))
            plt.plot(time, 5 * signal, label="Signal")
            plt.title("OFDM Frame")
            plt.xlabel("time")
            plt.legend()
            plt.savefig("OFDM Frame")
            plt.show()
        return signal


#########################################
#               Receiver                #
#########################################

class receiver(transmitter):

    def _sync(self, r):
        """Device matched filter + peak picking; returns (peak indices into `zeros`, len(zeros))."""
        import torch
        phy = self.phy
        r = np.asarray(r)
        if r.dtype in (np.uint8, np.int16):      # PCM as recorded: the narrow samples cross PCIe and are read by the kernels as they are
            d_r = torch.from_numpy(np.ascontiguousarray(r).reshape(1, -1)).to(phy.device)
        else:
            d_r = torch.from_numpy(np.ascontiguousarray(r, dtype=np.float32).reshape(1, -1)).to(phy.device)
        T = d_r.shape[1]
        max_peaks = max(4, T // max(1, phy.chirp_len) + 2)
        _, _, peaks, count = phy.sync_streams(d_r, max_peaks)
        n = int(count[0].item())
        return peaks[0, :n].cpu().numpy(), T + phy.chirp_len - 3, d_r

    def chirp_method(self, r):
        """OFDM.py:356-372: boolean `zeros` array with the surviving detections."""
        peaks, nz, _ = self._sync(r)
        zeros = np.zeros(nz, dtype=bool)
        zeros[peaks] = True
        return zeros

    def schmidlcox_method(self, r):
        """OFDM.py:376-387: Schmidl & Cox timing metric over the first 5 s (a prefix-sum kernel on the device) ->
        argmax |P| + N - 1.  Unused by receive() in the reference too ("no longer works with packets")."""
        import torch
        phy = self.phy
        r = np.asarray(r)
        dt = r.dtype if r.dtype in (np.uint8, np.int16) else np.float32
        d_r = torch.from_numpy(np.ascontiguousarray(r, dtype=dt).reshape(1, -1)).to(phy.device)
        search = 5 * int(self.fs)
        if d_r.shape[1] < search - 1 + 2 * self.L:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (search - 2 + 2 * self.L, d_r.shape[1]))
        idx, _ = phy.schmidlcox(d_r, search)
        return int(idx[0].item()) + self.ofdm_symbol_size - 1

    def get_symbols(self, r, zeros):
        """OFDM.py:391-403 (index bookkeeping on the host)."""
        zero_indicies = np.where(zeros == True)[0] + 2   # noqa: E712
        zero_indicies = zero_indicies[:-1]
        self.no_packets = len(zero_indicies)
        n = (2 * self.no_pilots + self.packet_length) * (self.cp_length + self.ofdm_symbol_size)
        rx = np.vstack([[r[i:i + n]] for i in zero_indicies])
        return rx.reshape(-1, 2 * self.no_pilots + self.packet_length, self.cp_length + self.ofdm_symbol_size)

    def remove_cp(self, rx):
        return rx[:, :, self.cp_length:]

    def get_data(self, OFDM_symbols):
        start_pilots = OFDM_symbols[:, :self.no_pilots, self.carriers]
        end_pilots = OFDM_symbols[:, -self.no_pilots:, self.carriers]
        data_symbols = OFDM_symbols[:, self.no_pilots:-self.no_pilots, self.carriers]
        return data_symbols, start_pilots, end_pilots

    def equalise(self, data_symbols, start_pilots, end_pilots):
        """OFDM.py:422-480 on spectra (what get_data returns), on the device:
        -> (data_eq[pk*L, K], Hest_start[pk, K], Hest_end[pk, K], Hest[pk, L, K])."""
        import torch
        data_symbols = np.asarray(data_symbols)
        if self.no_pilots == 0:
            return data_symbols.reshape(-1, self.K)                # OFDM.py:424-425 (different return arity)
        phy = self.phy
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex64)).to(phy.device)
        d_start, d_end, d_data = to_dev(start_pilots), to_dev(end_pilots), to_dev(data_symbols)
        Hs, He, slope = phy.eq_estimate(d_start, d_end)
        eq, hest = phy.eq_apply(d_data, Hs, He, slope)
        c128 = lambda t: t.cpu().numpy().astype(np.complex128)
        return c128(eq).reshape(-1, self.K), c128(Hs), c128(He), c128(hest)

    def demap(self, symbols):
        """OFDM.py:484-500: minimum-distance QPSK decisions -> (bits[n, carriers, 2], hardDecision)."""
        import torch
        if type(symbols) != np.ndarray:                            # noqa: E721  (the reference's own check, OFDM.py:485)
            raise ValueError("Symbols must be numpy array")
        phy = self.phy
        d = torch.from_numpy(np.ascontiguousarray(symbols, dtype=np.complex64)).to(phy.device)
        bits, hard = phy.demap(d)
        return bits.cpu().numpy().astype(np.int64), hard.cpu().numpy().astype(np.complex128)

    def PS(self, bits):
        return bits.reshape((-1,))

    def decode(self, bits_encoded):
        """OFDM.py:509-549 (receive() applies the XOR inside the demod kernel instead)."""
        if self.encoding == "LDPC":
            raise NotImplementedError('encoding "LDPC" is marked broken in the reference (OFDM.py:21) and is out of scope')
        if self.encoding == "XOR":
            dbs = self.data_bits_per_symbol
            known_bits = np.tile(self.known_sequence[:dbs], int(np.ceil(len(bits_encoded) / dbs)))[:len(bits_encoded)]
            return np.bitwise_xor(bits_encoded, known_bits)
        return bits_encoded

    def channel_response(self, Hest):
        """OFDM.py:553-577: plots of |H|, arg H and the impulse response of a channel estimate
        (Final System Test.ipynb calls it on receive()'s Hest_start; needs matplotlib)."""
        f = self.carriers / self.ofdm_symbol_size * self.fs
        for y, yl, name in ((abs(Hest), "|H(f)|", "plots/Channel_mag"), (np.angle(Hest), "arg(H(f))", "plots/Channel_freq")):
            plt.plot(f, y, label='Estimated channel')
            plt.ylabel(yl)
            plt.xlabel("Frequency")
            plt.title("Channel Frequency Response Estimate")
            plt.savefig(name)
            plt.show()
        h = np.fft.ifft(Hest)
        time = np.linspace(0, (len(h)), len(h))
        plt.plot(time[:500], h.real[:500])
        plt.title("Channel Impulse Response")
        plt.ylabel("h")
        plt.xlabel("time (samples)")
        plt.savefig("plots/Channel_inpulse")
        plt.show()

    def receive_packets(self, rx_cp, want_eq=False):
        """Device receive chain on already-sliced packets rx_cp[pk, 2P+L, N+cp] (what get_symbols
        returns): returns dict(bits, Hs, He, slope[, eq]).  Rows 8-12 of SURVEY 8a in two launches."""
        import torch
        phy = self.phy
        rx_cp = np.ascontiguousarray(np.asarray(rx_cp, dtype=np.float32))
        n_packets = rx_cp.shape[0]
        self.no_packets = n_packets
        d = torch.from_numpy(rx_cp.reshape(-1)).to(phy.device)
        return self._demod_device(d, n_packets, None, want_eq)

    def _demod_device(self, d_samples, n_packets, d_off, want_eq, packed=False):
        import torch
        phy = self.phy
        rx = phy.rx_receive if d_samples.dtype == torch.float32 else phy.rx_receive_pcm
        res, Hs, He, slope = rx(d_samples, n_packets, d_off, xor=(self.encoding == "XOR"), want_eq=want_eq)
        bits_packed, eq = res if want_eq else (res, None)
        if packed:                                                 # receive_bytes: the rows' bytes, no host bit array
            return dict(bits=bits_packed[:, : phy.bits_per_packet // 8].cpu().numpy().reshape(-1), bits_packed=bits_packed,
                        Hs=Hs.cpu().numpy().astype(np.complex128), He=He.cpu().numpy().astype(np.complex128), slope=slope.cpu().numpy())
        out = dict(bits=phy.unpack_bits(bits_packed), Hs=Hs.cpu().numpy().astype(np.complex128),
                   He=He.cpu().numpy().astype(np.complex128), slope=slope.cpu().numpy())
        if want_eq:
            out["eq"] = eq.cpu().numpy().reshape(-1, phy.K)
        return out

    def receive_bytes(self, signal):
        """np.packbits(receive(signal)[0]) without the host bit array: the packed rows the kernel wrote are
        the file's bytes.  Returns (bytes uint8, Hest_start[0], Hest_end[0]); save_file_bytes() writes them out and
        bit_errors() counts differences against a file on the device."""
        if (self.data_bits_per_symbol * self.packet_length) % 8:
            bits, hs, he = self.receive(signal)[:3] if not self.old_api else (self.receive(signal), None, None)
            return np.packbits(bits), hs, he
        return self.receive(signal, _packed=True)

    def bit_errors(self, a, b):
        """(bit errors, bits compared) between two packed byte strings over their common length, counted on the
        device (gf3_ber_count): the notebook's `np.sum(rx_bits[:n] != tx_bits) / n` (Final System Test.ipynb:150-160)."""
        import torch
        phy = self.phy
        a = np.ascontiguousarray(np.asarray(a, dtype=np.uint8).reshape(-1)); b = np.ascontiguousarray(np.asarray(b, dtype=np.uint8).reshape(-1))
        n = min(len(a), len(b))
        if n == 0:
            return 0, 0
        da = torch.from_numpy(a[:n]).to(phy.device); db = torch.from_numpy(b[:n]).to(phy.device)
        counter = torch.zeros(2, dtype=torch.int64, device=phy.device)
        phy.ber_count(da, db, 8 * n, counter)
        c = counter.cpu().numpy()
        return int(c[0]), int(c[1])

    def receive(self, signal, graph_output=False, _details=None, _packed=False):
        """OFDM.py:581-657: sync -> slice -> FFT -> equalise -> demap -> decode, on the device."""
        import torch
        print("-" * 42 + "\nReceive \n" + "-" * 42)
        print("OFDM Paramters:")
        print(self)
        if self.encoding == "LDPC":
            raise NotImplementedError('encoding "LDPC" is out of scope (broken in the reference, OFDM.py:21)')
        if self.sync_method != "chirp":
            raise NotImplementedError("sync_method %r: only the chirp synchronisation is built (OFDM.py:56 hard-wires it; the reference "
                                      "marks schmidlcox_method as no longer working, OFDM.py:376)" % (self.sync_method,))
        phy = self.phy
        peaks, nz, d_r = self._sync(signal)
        starts = (peaks + 2)[:-1]                                  # OFDM.py:393-395
        self.no_packets = len(starts)
        if self.no_packets == 0:
            raise ValueError("need at least one array to concatenate")            # np.vstack([]) in OFDM.py:400
        T = d_r.shape[1]
        if starts[-1] + phy.pkt_samples > T:                       # ragged slices -> vstack/reshape error in the reference
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly "
                             "(signal ends inside the last packet)")
        d_off = torch.from_numpy(starts.astype(np.int64)).to(phy.device)
        print("Number of received OFDM symbols:    " + str(self.no_packets * self.packet_length))
        out = self._demod_device(d_r.reshape(-1), self.no_packets, d_off, want_eq=_details is not None, packed=_packed)
        bits = out["bits"]
        print("Number of received bits:            " + str(len(bits) * (8 if _packed else 1)))
        if _details is not None:
            _details.update(out, peaks=peaks, starts=starts)
        self.Hest = out["Hs"][0]                                   # old API attribute (Week 2 Challenge.ipynb:305)
        if self.old_api:
            return bits                                            # the old receive() returned the bits alone (Audio.ipynb:147-161)
        return bits, out["Hs"][0], out["He"][0]


def play_record(signal, fs, padding_before=1, padding_after=1):
    """OFDM.py:681-689 (needs the optional sounddevice module)."""
    data_padded = np.pad(signal, (int(padding_before * fs), int(padding_after * fs)), 'constant', constant_values=0)
    print("Recording...")
    signal = sd.playrec(data_padded, fs)
    sd.wait()
    print("Finished recording")
    return signal[:, 0]


class channel(receiver):
    """OFDM.py:694-752 is dead code in the reference (calls a nonexistent self.pad); kept as a name."""

    def measure_channel(self, bits):
        raise NotImplementedError("channel.measure_channel is dead code in the reference (OFDM.py:707 calls self.pad)")


def load_file(file_name):
    """OFDM.py:756-761."""
    data_bytes = np.fromfile(_find_ci(os.path.join("input_files", file_name)), dtype=np.uint8)
    file_info = file_name + "\x00" + str(len(data_bytes)) + "\x00"
    b = bytearray()
    b.extend(map(ord, file_info))
    return np.unpackbits(np.hstack([b, data_bytes]))


def load_file_bytes(file_name):
    """np.packbits(load_file(file_name)): header + data as bytes, for transmitter.transmit_bytes."""
    data_bytes = np.fromfile(_find_ci(os.path.join("input_files", file_name)), dtype=np.uint8)
    file_info = file_name + "\x00" + str(len(data_bytes)) + "\x00"
    return np.hstack([np.frombuffer(file_info.encode("latin-1"), dtype=np.uint8), data_bytes])


def save_file_bytes(rx_bytes):
    """save_file(np.unpackbits(rx_bytes)): for receiver.receive_bytes."""
    return save_file(rx_bytes, _packed=True)


def save_file(rx_bits, _packed=False):
    """OFDM.py:766-794."""
    data = np.asarray(rx_bits, dtype=np.uint8) if _packed else np.packbits(rx_bits)
    z1 = int(np.flatnonzero(data == 0)[0])
    file_name = "".join(chr(c) for c in data[:z1])
    data = data[z1 + 1:]
    z2 = int(np.flatnonzero(data == 0)[0])
    file_size = "".join(chr(c) for c in data[:z2])
    data = data[z2 + 1:]
    print("File Name: " + file_name + "\nFile Size: " + file_size + " bytes")
    data = data[:int(file_size)]
    out_dir = _find_ci("output_files")
    data.tofile(os.path.join(out_dir, file_name[:-4] + "_received" + file_name[-4:]))
    return file_name, data
