"""CPU oracle for the GF3 OFDM physical-layer hot path (float64 numpy).

TEST INFRASTRUCTURE ONLY -- never imported by the product (gf3-audio-modem_b200/).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it,
and only as the checker or the timed CPU arm.

This is a vectorised restatement of the reference's algorithm; every function cites the
reference lines (file:line into /root/reference) it follows.  The arithmetic is the reference's
(float64 / complex128, same numpy / scipy library calls where the reference makes them:
np.fft.fft / ifft, scipy.signal.convolve, scipy.signal.chirp, np.unwrap, np.polyfit); only the
Python loops over packets / symbols / carriers were replaced by array expressions.

Parity status: PINNED.  tests/test_oracle_golden.py checks it against golden vectors generated
by running the unmodified reference in the build container (oracle/make_golden.py, fixtures in
tests/golden/), including the reference's one published known answer: BER
0.023375665289067146 on received_signals/gr5ch1_signal.wav (Final System Test.ipynb:160).
tests/test_oracle_vs_reference.py re-runs the comparison live whenever /root/reference exists.
"""
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "Params", "MODES", "load_known_sequence", "sync_chirp", "qpsk_map", "known_symbols",
    "encode", "decode", "random_qpsk", "build_ofdm_symbols", "add_cp", "send_to_stream",
    "transmit", "matched_filter", "chirp_method", "schmidlcox_method", "get_symbols", "rx_fft", "get_data",
    "equalise", "demap", "receive", "receive_symbols", "known_channel_decode",
    "load_file_bits", "save_file_bytes",
]

# OFDM.py:30-40
MODES = {
    "A1": (224, (1, 2047)), "A2": (224, (100, 1500)), "A3": (224, (100, 1000)),
    "B1": (704, (1, 2047)), "B2": (704, (100, 1500)), "B3": (704, (100, 1000)),
    "C1": (1184, (1, 2047)), "C2": (1184, (100, 1500)), "C3": (1184, (100, 1000)),
}


@dataclass
class Params:
    """Parameter contract of CamG.__init__ (OFDM.py:18-101), generalised to any (N, cp, lo, hi)."""
    N: int = 4096            # OFDM.py:27
    cp: int = 224            # OFDM.py:42
    lo: int = 100            # OFDM.py:43  lowest data bin (inclusive)
    hi: int = 1500           # OFDM.py:44  highest data bin (EXCLUSIVE: np.arange(lo, hi), OFDM.py:47)
    n_pilots: int = 20       # OFDM.py:51
    packet_len: int = 180    # OFDM.py:50
    fs: int = 48000          # OFDM.py:24
    f0: float = 0.0          # OFDM.py:62
    f1: float = 8000.0       # OFDM.py:63
    encoding: str = "XOR"    # OFDM.py:21
    thresh: float = 0.4      # OFDM.py:361
    fit_lo: int = 500        # OFDM.py:462
    fit_hi: int = 1000       # OFDM.py:462
    known_sequence: np.ndarray = field(default=None, repr=False)  # OFDM.py:99-101

    @classmethod
    def from_mode(cls, mode, **kw):
        cp, (lo, hi) = MODES[mode]
        return cls(N=4096, cp=cp, lo=lo, hi=hi, **kw)

    # derived attributes, OFDM.py:28,46-49,64,94-95
    @property
    def K(self):
        return self.N // 2 - 1

    @property
    def carriers(self):
        return np.arange(1, self.K + 1)

    @property
    def data_carriers(self):
        return np.arange(self.lo, self.hi)

    @property
    def Nd(self):
        return self.hi - self.lo

    @property
    def unused_carriers(self):
        return np.delete(self.carriers, self.data_carriers - 1)

    @property
    def chirp_length(self):
        return 5 * (self.N + self.cp)

    @property
    def data_bits_per_symbol(self):
        return 2 * self.Nd

    @property
    def bits_per_symbol(self):
        return 2 * self.K

    @property
    def sym_len(self):
        return self.N + self.cp

    @property
    def syms_per_packet(self):
        return 2 * self.n_pilots + self.packet_len

    @property
    def packet_samples(self):
        return self.syms_per_packet * self.sym_len


def load_known_sequence(path, n=4096):
    """First n characters of Handouts/random_bits.txt as 0/1 ints (OFDM.py:99-101)."""
    with open(path, "rb") as f:
        raw = f.read(n)
    return (np.frombuffer(raw, dtype=np.uint8) - ord("0")).astype(np.int64)


# ----------------------------------------------------------------------------- sync chirp
def sync_chirp(p):
    """OFDM.py:106-109 -- linear chirp f0->f1 over Lc samples, endpoint-inclusive time base, /5.

    scipy.signal.chirp(method='linear') is cos(2*pi*(f0*t + 0.5*(f1-f0)/t1*t^2)); restated in
    closed form (verified to 1e-12 against scipy in tests).
    """
    Lc = p.chirp_length
    t1 = Lc / p.fs
    t = np.linspace(0, t1, Lc)
    beta = (p.f1 - p.f0) / t1
    return np.cos(2 * np.pi * (p.f0 * t + 0.5 * beta * t * t)) / 5


# ----------------------------------------------------------------------------- QPSK
def qpsk_map(bits2):
    """OFDM.py:72-77,196-197 -- Gray QPSK: (b0,b1) -> ((1-2*b1) + 1j*(1-2*b0))/sqrt(2).

    (0,0)->(1+1j), (1,0)->(1-1j), (1,1)->(-1-1j), (0,1)->(-1+1j), all /sqrt(2).
    bits2[..., 2] -> complex128[...]
    """
    b = np.asarray(bits2).astype(np.int64)          # uint8 input would wrap in 1-2*b
    # the reference stores the table entries as (x+yj)/np.sqrt(2): reproduce that division
    return ((1 - 2 * b[..., 1]) + 1j * (1 - 2 * b[..., 0])) / np.sqrt(2)


def known_symbols(p):
    """OFDM.py:244,429 -- the known (pilot) OFDM symbol: first 2K known bits mapped on all K bins."""
    return qpsk_map(p.known_sequence[: p.bits_per_symbol].reshape(p.K, 2))


# ----------------------------------------------------------------------------- bit coding
def encode(p, bits, rng=np.random):
    """OFDM.py:163-185 -- XOR with tiled known bits (encoding "XOR"), pad to whole packets with
    Bernoulli(1/2) draws from the numpy global RNG (rng.binomial), same call order."""
    bits = np.asarray(bits)
    if p.encoding == "XOR":
        dbs = p.data_bits_per_symbol
        known = np.tile(p.known_sequence[:dbs], int(np.ceil(len(bits) / dbs)))[: len(bits)]
        bits = np.bitwise_xor(bits, known)
    bpp = p.data_bits_per_symbol * p.packet_len
    pad_len = (bpp - len(bits) % bpp) % bpp
    padding = rng.binomial(n=1, p=0.5, size=(pad_len,))
    return np.hstack([bits, padding])


def decode(p, bits_encoded):
    """OFDM.py:541-547."""
    if p.encoding == "XOR":
        dbs = p.data_bits_per_symbol
        known = np.tile(p.known_sequence[:dbs], int(np.ceil(len(bits_encoded) / dbs)))[: len(bits_encoded)]
        return np.bitwise_xor(bits_encoded, known)
    return bits_encoded


# ----------------------------------------------------------------------------- transmit chain
def random_qpsk(p, rng=np.random):
    """OFDM.py:201-203 -- ONE filler vector per transmit() call, drawn with rng.choice."""
    qpsk = np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
    return rng.choice(qpsk, size=(p.K - p.Nd), replace=True)


def build_ofdm_symbols(p, payload, filler):
    """OFDM.py:207-217 -- Hermitian-symmetric N-bin spectrum; DC and Nyquist stay 0."""
    X = np.zeros([payload.shape[0], p.N], dtype=complex)
    X[:, p.data_carriers] = payload
    X[:, p.unused_carriers] = filler
    X[:, -p.data_carriers] = np.conj(payload)
    X[:, -p.unused_carriers] = np.conj(filler)
    return X


def add_cp(p, time_data):
    """OFDM.py:221-226."""
    if p.cp == 0:
        return time_data
    return np.hstack([time_data[:, -p.cp:], time_data])


def send_to_stream(p, time_data_cp, sync):
    """OFDM.py:244-259 -- [chirp | 2*(P x known) | 2*(L x data) | 2*(P x known)] per packet, plus
    one trailing chirp.  (The plot masks of :262-273 are out of scope.)"""
    ks = known_symbols(p).reshape(1, p.K)
    known_ofdm = np.zeros([1, p.N], dtype=complex)
    known_ofdm[0, p.carriers] = ks
    known_ofdm[0, -p.carriers] = np.conj(ks)
    known_time = add_cp(p, np.fft.ifft(known_ofdm))
    packets = time_data_cp.reshape(-1, p.packet_len, p.sym_len)
    n_packets = packets.shape[0]
    known_time = np.tile(known_time, (n_packets, p.n_pilots, 1))
    sync_t = np.tile(sync, (n_packets, 1))
    tx = 2 * np.hstack([known_time, packets, known_time])
    tx = np.hstack([sync_t, tx.reshape(n_packets, -1)])
    tx = tx.reshape(-1).real
    return np.hstack([tx, sync_t[0]]), n_packets


def transmit(p, bits, rng=np.random):
    """OFDM.py:296-343 -- encode -> SP -> map -> build -> ifft -> add_cp -> send_to_stream.
    RNG draw order: binomial padding (encode) then choice filler (build_OFDM_symbol)."""
    enc = encode(p, bits, rng)
    payload = qpsk_map(enc.reshape(-1, p.Nd, 2))           # SP + map, OFDM.py:191-197
    filler = random_qpsk(p, rng)
    X = build_ofdm_symbols(p, payload, filler)
    x = add_cp(p, np.fft.ifft(X))                          # OFDM.py:322-323
    tx, _ = send_to_stream(p, x, sync_chirp(p))
    return tx


# ----------------------------------------------------------------------------- synchronisation
def matched_filter(p, r):
    """OFDM.py:357-358 -- full linear convolution with the time-reversed chirp (same scipy call)."""
    from scipy.signal import convolve
    return convolve(r, sync_chirp(p)[::-1], mode="full")


def chirp_method(p, r, P=None):
    """OFDM.py:356-372 -- normalise by the (signed) global max, candidate mask
    (D[i]*D[i+1] <= 0) & (P[i+1] > thresh), then the ascending hold-off scan: every surviving
    candidate clears the next Lc entries.  Quirk (OFDM.py:366-370): if the clearing runs off the
    end (IndexError) the `except` wipes zeros[:i+1], so NO detection survives at all."""
    if P is None:
        P = matched_filter(p, r)
    P = P / np.amax(P)
    D = np.diff(P)
    zeros = ((D[:-1] * D[1:]) <= 0) & (P[1:-1] > p.thresh)
    Lc = p.chirp_length
    n = len(zeros)
    cand = np.flatnonzero(zeros)
    out = np.zeros(n, dtype=bool)
    next_free = 0
    for idx in cand:                      # loop over candidates only (a few per chirp)
        if idx < next_free:
            continue
        if idx + Lc >= n:                 # zeros[idx+1+j] would raise for some j < Lc
            out[:] = False
            return out
        out[idx] = True
        next_free = idx + Lc + 1
    return out


def schmidlcox_method(p, r):
    """OFDM.py:376-387 -- Schmidl & Cox timing metric over the first 5 s: the recursion
    P[d+1] = P[d] + r[d+L] r[d+2L] - r[d] r[d+L] (L = K + 1 = N/2, OFDM.py:54) is a cumulative sum (np.cumsum adds in
    the same sequential order as the reference's loop); returns argmax |P| + N - 1 (first occurrence)."""
    r = np.asarray(r, dtype=np.float64)
    n = 5 * p.fs
    L = p.N // 2
    d = np.arange(n - 1)
    terms = r[d + L] * r[d + 2 * L] - r[d] * r[d + L]
    P = np.concatenate([[0.0], np.cumsum(terms)])
    return int(np.where(np.abs(P) == np.amax(np.abs(P)))[0][0]) + p.N - 1


def get_symbols(p, r, zeros):
    """OFDM.py:391-403 -- packet starts = where(zeros)+2, last detection dropped, slice
    (2P+L)(N+cp) samples per packet.  Returns (rx[pk, 2P+L, N+cp], starts)."""
    starts = np.where(zeros)[0] + 2
    starts = starts[:-1]
    n = p.packet_samples
    rx = np.vstack([[r[i:i + n]] for i in starts])
    return rx.reshape(-1, p.syms_per_packet, p.sym_len), starts


# ----------------------------------------------------------------------------- receive chain
def rx_fft(p, rx_cp):
    """OFDM.py:407-408,593 -- strip the CP, N-point DFT (complex128, unnormalised)."""
    return np.fft.fft(rx_cp[:, :, p.cp:])


def get_data(p, ofdm):
    """OFDM.py:412-418 -- bins 1..K; pilots [:P] / data [P:-P] / pilots [-P:]."""
    P = p.n_pilots
    c = p.carriers
    return ofdm[:, P:-P, c], ofdm[:, :P, c], ofdm[:, -P:, c]


def equalise(p, data, start_pilots, end_pilots):
    """OFDM.py:422-480.

    Hs = mean_P(start)/known, He = mean_P(end)/known                      (:443-451)
    dphi = unwrap(angle(He)) - unwrap(angle(Hs)) along bins               (:454-457)
    slope = polyfit(arange(len(w)), dphi[fit_lo:fit_hi], 1)[0]            (:461-462)
    for symbol l, 0-based carrier index n:  w = (l + P/2)/(L + P)         (:471,474)
      |H| = |Hs| + (|He|-|Hs|) w ;  theta = angle(Hs) + slope*n*w ;  out = data/(|H| e^{j theta})
    Returns (data_eq[pk*L, K], Hs[pk,K], He[pk,K], Hest[pk,L,K], slope[pk]).
    """
    P, L, K = p.n_pilots, p.packet_len, p.K
    if P == 0:
        return data.reshape(-1, K), None, None, None, None
    ks = known_symbols(p)
    Hs = (np.mean(start_pilots.real, axis=1) + 1j * np.mean(start_pilots.imag, axis=1)) / ks
    He = (np.mean(end_pilots.real, axis=1) + 1j * np.mean(end_pilots.imag, axis=1)) / ks
    dphi = np.unwrap(np.angle(He)) - np.unwrap(np.angle(Hs))
    npk = data.shape[0]
    slope = np.zeros(npk)
    for i in range(npk):
        win = dphi[i, p.fit_lo:p.fit_hi]
        slope[i] = np.polyfit(np.arange(len(win)), win, 1)[0]
    w = ((np.arange(L) + P / 2) / (L + P))[None, :, None]
    n = np.arange(K)[None, None, :]
    mag = np.abs(Hs)[:, None, :] + (np.abs(He) - np.abs(Hs))[:, None, :] * w
    theta = np.angle(Hs)[:, None, :] + slope[:, None, None] * n * w
    Hest = mag * np.exp(1j * theta)
    return (data / Hest).reshape(-1, K), Hs, He, Hest, slope


def demap(symbols):
    """OFDM.py:484-500 -- arg-min distance over [(0,0),(1,0),(1,1),(0,1)] (dict order).  Equivalent
    to b0 = imag<0, b1 = real<0; on exact ties (a component == 0.0) argmin
    keeps the FIRST minimum, which gives b1 = 0 always and b0 = 1 only for imag == 0 with
    real < 0 (pinned by tests/test_oracle_golden.py::test_demap_equals_min_distance).
    Returns int64 bits[..., 2]."""
    s = np.asarray(symbols)
    b0 = (s.imag < 0) | ((s.imag == 0) & (s.real < 0))
    return np.stack([b0, (s.real < 0)], axis=-1).astype(np.int64)


def demap_min_distance(symbols):
    """Literal form of OFDM.py:487-500 (used by the tests to pin `demap` above)."""
    pts = np.array([(1 + 1j), (1 - 1j), (-1 - 1j), (-1 + 1j)]) / np.sqrt(2)
    lab = np.array([(0, 0), (1, 0), (1, 1), (0, 1)])
    d = np.abs(np.asarray(symbols)[..., None] - pts)
    return lab[d.argmin(axis=-1)]


def receive_symbols(p, rx_cp, want_eq=False):
    """OFDM.py:591-609 for already-sliced packets rx_cp[pk, 2P+L, N+cp] (rows 8-12 of SURVEY 8a).
    Returns dict(bits, Hs, He, slope[, eq])."""
    data, sp, ep = get_data(p, rx_fft(p, rx_cp))
    eq, Hs, He, _, slope = equalise(p, data, sp, ep)
    eq_d = eq[:, p.data_carriers - 1]                       # OFDM.py:603
    bits = decode(p, demap(eq_d).reshape(-1))              # OFDM.py:605-609
    out = dict(bits=bits, Hs=Hs, He=He, slope=slope)
    if want_eq:
        out["eq"] = eq
    return out


def receive(p, signal, want_eq=False):
    """OFDM.py:581-657 -- chirp_method -> get_symbols -> receive_symbols."""
    zeros = chirp_method(p, signal)
    rx_cp, starts = get_symbols(p, signal, zeros)
    out = receive_symbols(p, rx_cp, want_eq)
    out["starts"] = starts
    out["peaks"] = np.where(zeros)[0]
    return out


def known_channel_decode(p, rx_cp, H):
    """Weekend Challenge.ipynb:162-226 (old API; code not in the repo): Y/H on bins 1..K, demap.
    rx_cp[n_sym, N+cp]; H = fft(h, N).  Returns (bits[n_sym*K*2], eq[n_sym, K])."""
    Y = np.fft.fft(rx_cp[:, p.cp:])
    eq = (Y / H)[:, 1:p.K + 1]
    return demap(eq).reshape(-1), eq


# ----------------------------------------------------------------------------- file framing
def load_file_bits(name, data_bytes):
    """OFDM.py:756-761 -- name\\0size\\0 header + payload, MSB-first bits."""
    info = (name + "\x00" + str(len(data_bytes)) + "\x00").encode("latin-1")
    return np.unpackbits(np.hstack([np.frombuffer(info, np.uint8), np.asarray(data_bytes, np.uint8)]))


def save_file_bytes(rx_bits):
    """OFDM.py:766-794 without the file write -- returns (name, size_str, payload bytes)."""
    data = np.packbits(rx_bits)
    z1 = int(np.flatnonzero(data == 0)[0])
    name = "".join(chr(c) for c in data[:z1])
    rest = data[z1 + 1:]
    z2 = int(np.flatnonzero(rest == 0)[0])
    size = "".join(chr(c) for c in rest[:z2])
    return name, size, rest[z2 + 1:][: int(size)]
