"""GPU parity tests: the CUDA path (through the C-ABI, libgf3b200.so) against the numpy oracle and
the committed reference goldens.  Bits must be identical except where the oracle's float64
constellation point lies within BOUNDARY_TOL of a QPSK decision boundary (those are counted and
reported separately, per the north star); equalised points within EQ_RTOL (fp32 vs float64)."""
import numpy as np
import pytest

from conftest import BOUNDARY_TOL, STAGE_NAMES, assert_bits_match, load_golden, oracle_params
from oracle import gf3_oracle as orc

pytestmark = pytest.mark.gpu

EQ_RTOL = 1e-4          # north star: equalised constellation within 1e-4 relative error


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _phy(p):
    import gf3b200
    return gf3b200.Phy(N=p.N, cp=p.cp, lo=p.lo, hi=p.hi, n_pilots=p.n_pilots, packet_len=p.packet_len,
                       known_sequence=p.known_sequence, fit_lo=p.fit_lo, fit_hi=p.fit_hi)


def _check_bits(got, ref_bits, ref_eq_data, what):
    """Bit-exact, except decisions whose oracle point is within 1e-5 of a boundary (conftest policy)."""
    return assert_bits_match(got, ref_bits, ref_eq_data, what)["n_diff"]


def _rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


# ----------------------------------------------------------------------------- FFT front end
@pytest.mark.parametrize("N,cp", [(64, 16), (128, 0), (256, 16), (512, 64), (1024, 32), (2048, 64), (4096, 224), (4096, 704),
                                  (1024, 33), (2048, 7)])                  # odd symbol length: every other symbol 4-byte aligned only
def test_spectrum_matches_numpy_fft(N, cp, known_sequence):
    torch = _torch()
    p = orc.Params(N=N, cp=cp, lo=1, hi=N // 2, n_pilots=0, packet_len=1, known_sequence=known_sequence)
    phy = _phy(p)
    rng = np.random.default_rng(N + cp)
    nsym = 37
    x = rng.normal(size=(nsym, N + cp)).astype(np.float32)
    ref = np.fft.fft(x[:, cp:].astype(np.float64))[:, 1:N // 2]
    got = phy.spectrum(torch.from_numpy(x).cuda().reshape(-1), nsym).cpu().numpy()
    scale = np.sqrt(np.mean(np.abs(ref) ** 2))
    assert np.max(np.abs(got - ref)) / scale < 3e-6
    # arbitrary (odd) sample offsets take the scalar-load path
    flat = np.concatenate([np.zeros(3, np.float32), x.reshape(-1)])
    offs = torch.from_numpy(3 + np.arange(nsym, dtype=np.int64) * (N + cp)).cuda()
    got2 = phy.spectrum(torch.from_numpy(flat).cuda(), nsym, offs).cpu().numpy()
    assert np.array_equal(got, got2)


# ----------------------------------------------------------------------------- stage goldens
@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_receive_chain(name, known_sequence):
    """rows 7-12 of SURVEY 8a on the reference's own stage outputs (tests/golden/stage_*.npz)."""
    torch = _torch()
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    phy = _phy(p)
    r = g["r_i16"].astype(np.float32)
    starts = (g["peaks"] + 2)[:-1]
    npk = len(starts)
    d = torch.from_numpy(r).cuda()
    off = torch.from_numpy(starts.astype(np.int64)).cuda()
    Hs, He, slope = phy.rx_estimate(d, npk, off)
    packed, eq = phy.rx_demod(d, npk, Hs, He, slope, off, xor=True, want_eq=True)
    Hs, He, slope, eq = Hs.cpu().numpy(), He.cpu().numpy(), slope.cpu().numpy(), eq.cpu().numpy().reshape(-1, p.K)
    hscale = np.max(np.abs(g["Hs"]))
    assert np.max(np.abs(Hs - g["Hs"])) / hscale < 2e-6
    assert np.max(np.abs(He - g["He"])) / hscale < 2e-6
    np.testing.assert_allclose(slope, g["slope"], rtol=0, atol=2e-7)
    # equalised constellation: data carriers within EQ_RTOL
    dc = p.data_carriers - 1
    err = _rel_err(eq[:, dc], g["eq"][:, dc])
    assert err.max() < EQ_RTOL, "max rel eq error %.3e" % err.max()
    _check_bits(phy.unpack_bits(packed), g["bits"], g["eq"][:, dc], name)
    # without the fused XOR the raw demapped bits must match too
    packed_raw = phy.rx_demod(d, npk, torch.from_numpy(Hs).cuda(), torch.from_numpy(He).cuda(),
                              torch.from_numpy(slope).cuda(), off, xor=False)
    _check_bits(phy.unpack_bits(packed_raw), g["bits_raw"], g["eq"][:, dc], name + " raw")
    # the whole chain in one launch (estimate fused into the data-symbol kernel): same bars
    (packed_f, eq_f), Hs_f, He_f, slope_f = phy.rx_receive(d, npk, off, xor=True, want_eq=True)
    assert np.max(np.abs(Hs_f.cpu().numpy() - g["Hs"])) / hscale < 2e-6
    assert np.max(np.abs(He_f.cpu().numpy() - g["He"])) / hscale < 2e-6
    np.testing.assert_allclose(slope_f.cpu().numpy(), g["slope"], rtol=0, atol=2e-7)
    assert _rel_err(eq_f.cpu().numpy().reshape(-1, p.K)[:, dc], g["eq"][:, dc]).max() < EQ_RTOL
    _check_bits(phy.unpack_bits(packed_f), g["bits"], g["eq"][:, dc], name + " fused")
    packed_f2, _, _, _ = phy.rx_receive(d, npk, off, xor=True)
    assert torch.equal(packed_f2, packed_f)


@pytest.mark.parametrize("dtype", ["int16", "float32"])
@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_receive_chain_staged_input(name, dtype, known_sequence):
    """The same goldens through gf3_rx_receive_pcm: the int16 recording read as int16 (converted in registers)
    and the float32 copy through the same cp.async.bulk staging buffer, packets at the reference's own
    (arbitrary, mostly unaligned) sync offsets -- same bars as the direct path."""
    torch = _torch()
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    phy = _phy(p)
    starts = (g["peaks"] + 2)[:-1]
    npk = len(starts)
    d = torch.from_numpy(g["r_i16"].astype(np.int16 if dtype == "int16" else np.float32)).cuda()
    off = torch.from_numpy(starts.astype(np.int64)).cuda()
    (packed, eq), Hs, He, slope = phy.rx_receive_pcm(d, npk, off, xor=True, want_eq=True)
    hscale = np.max(np.abs(g["Hs"]))
    assert np.max(np.abs(Hs.cpu().numpy() - g["Hs"])) / hscale < 2e-6
    assert np.max(np.abs(He.cpu().numpy() - g["He"])) / hscale < 2e-6
    np.testing.assert_allclose(slope.cpu().numpy(), g["slope"], rtol=0, atol=2e-7)
    dc = p.data_carriers - 1
    assert _rel_err(eq.cpu().numpy().reshape(-1, p.K)[:, dc], g["eq"][:, dc]).max() < EQ_RTOL
    _check_bits(phy.unpack_bits(packed), g["bits"], g["eq"][:, dc], "%s staged %s" % (name, dtype))
    packed2, _, _, _ = phy.rx_receive_pcm(d, npk, off, xor=True)            # throughput variant (no constellation output)
    _check_bits(phy.unpack_bits(packed2), g["bits"], g["eq"][:, dc], "%s staged %s, bits only" % (name, dtype))
    # shifted copies of the recording: every 16-byte misalignment of the bulk copies' source
    for shift in (1, 2, 3, 5):
        buf = torch.zeros(d.numel() + 8, dtype=d.dtype, device="cuda")
        buf[shift:shift + d.numel()] = d
        p3, _, _, _ = phy.rx_receive_pcm(buf, npk, off + shift, xor=True)
        assert torch.equal(p3, packed2), shift


def test_uint8_receive_at_scale_vs_oracle(known_sequence):
    """C3 shape, 192 packets quantised to 8-bit PCM with the wav offset of 128 (the reference's recording format,
    Final System Test.ipynb:85-86): gf3_rx_receive_pcm on the uint8 samples against the oracle on `r/1.0`."""
    torch = _torch()
    import gf3b200
    from gf3b200 import synth
    n = 192
    cfg = dict(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, fit_lo=125, fit_hi=250)
    phy = gf3b200.Phy(known_sequence=known_sequence, **cfg)
    p = orc.Params(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, known_sequence=known_sequence, encoding="XOR", fit_lo=125, fit_hi=250)
    b = synth.make_batch(phy, n, 1, snr_db=14.0, seed=31)
    sym = synth.packets_from_streams(phy, b)
    q = (torch.round(sym * (110.0 / float(sym.abs().max()))) + 128.0).to(torch.uint8).contiguous()
    (packed, eq), Hs, He, slope = phy.rx_receive_pcm(q.reshape(-1), n, xor=True, want_eq=True)
    fast, _, _, _ = phy.rx_receive_pcm(q.reshape(-1), n, xor=True)
    ref = orc.receive_symbols(p, q.cpu().numpy().astype(np.float64).reshape(n, p.syms_per_packet, p.sym_len), want_eq=True)
    hscale = np.max(np.abs(ref["Hs"]), axis=1, keepdims=True)
    assert np.max(np.abs(Hs.cpu().numpy() - ref["Hs"]) / hscale) < 2e-6
    np.testing.assert_allclose(slope.cpu().numpy(), ref["slope"], rtol=0, atol=2e-7)
    dc = p.data_carriers - 1
    ref_eq = ref["eq"][:, dc]
    rel = np.abs(eq.cpu().numpy().reshape(-1, p.K)[:, dc] - ref_eq) / np.maximum(np.abs(ref_eq), 1.0)
    assert np.quantile(rel, 0.9999) < EQ_RTOL
    _check_bits(phy.unpack_bits(packed), ref["bits"], ref_eq, "uint8 C3")
    _check_bits(phy.unpack_bits(fast), ref["bits"], ref_eq, "uint8 C3, bits only")


@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_level_methods(name, known_sequence):
    """receiver.equalise / receiver.demap / transmitter.send_to_stream as separate public methods
    (OFDM.py:422-480, 484-500, 242-276), fed with the oracle's spectra of the golden recordings."""
    import OFDM
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    N, cp, lo, hi, P, L = (int(x) for x in g["cfg"][:6])
    rx = OFDM.receiver("A2", "XOR", no_pilots=P, packet_length=L, ofdm_symbol_size=N, cp_length=cp,
                       lowest_bin=lo, highest_bin=hi)
    rx.known_sequence = np.asarray(known_sequence)
    r = g["r_i16"].astype(np.float64)
    starts = (g["peaks"] + 2)[:-1]
    n = (2 * P + L) * (N + cp)
    rx_cp = np.stack([r[s:s + n] for s in starts]).reshape(len(starts), 2 * P + L, N + cp)
    data, sp, ep = orc.get_data(p, orc.rx_fft(p, rx_cp))
    rx.no_packets = len(starts)
    eq, Hs, He, Hest = rx.equalise(data, sp, ep)
    ref_eq, rHs, rHe, rHest, _ = orc.equalise(p, data, sp, ep)
    assert eq.shape == ref_eq.shape and Hest.shape == rHest.shape and eq.dtype == np.complex128
    hscale = np.max(np.abs(rHs))
    assert np.max(np.abs(Hs - rHs)) / hscale < 2e-6 and np.max(np.abs(He - rHe)) / hscale < 2e-6
    assert np.max(np.abs(Hs - g["Hs"])) / hscale < 2e-6            # and the reference's own values
    dc = p.data_carriers - 1
    assert _rel_err(eq[:, dc], ref_eq[:, dc]).max() < EQ_RTOL
    assert _rel_err(Hest[:, :, dc], rHest[:, :, dc]).max() < EQ_RTOL
    # demap: same decisions as the oracle on the oracle's own constellation, exact ties included
    pts = np.ascontiguousarray(ref_eq[:, dc])
    bits, hard = rx.demap(pts)
    assert bits.dtype == np.int64 and bits.shape == pts.shape + (2,)
    assert np.array_equal(bits, orc.demap(pts.astype(np.complex64)))
    assert np.allclose(hard, ((1 - 2 * bits[..., 1]) + 1j * (1 - 2 * bits[..., 0])) / np.sqrt(2))
    ties = np.array([[0, 1, -1, 1j, -1j, 1 + 1j, -1 - 1j, -0.0 + 2j, 3 - 0.0j, -3 - 0.0j]], dtype=complex)
    tb, _ = rx.demap(ties)
    assert np.array_equal(tb, orc.demap_min_distance(ties))
    with pytest.raises(ValueError):
        rx.demap([1 + 1j])
    # send_to_stream: framing of time-domain symbols with the caller's sync waveform
    rng = np.random.default_rng(3)
    npk = 2
    time_data = rng.standard_normal((npk * L, N + cp)) * 0.05
    sync = orc.sync_chirp(p)
    tx, sv, kv, pv = rx.send_to_stream(time_data, sync)
    ref_tx, ref_npk = orc.send_to_stream(p, time_data, sync)
    assert rx.no_packets == ref_npk == npk and tx.shape == ref_tx.shape
    assert np.max(np.abs(tx - ref_tx)) < 3e-7 * max(1.0, np.max(np.abs(ref_tx)))
    plen = len(sync) + (2 * P + L) * (N + cp)
    assert sv.shape == kv.shape == pv.shape == (npk * (npk * plen + len(sync)),)   # OFDM.py:274 tiles the whole frame
    assert kv.sum() == npk * 2 * P * N and pv.sum() == npk * L * N and sv.sum() == npk * 2 * len(sync)


@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_sync(name, known_sequence):
    """row 6: matched filter + detection rule give the reference's sync indices."""
    torch = _torch()
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    phy = _phy(p)
    r = g["r_i16"].astype(np.float32).reshape(1, -1)
    d = torch.from_numpy(r).cuda()
    P, pmax = phy.xcorr(d)
    Pref = orc.matched_filter(p, r[0].astype(np.float64))
    assert P.shape[1] == len(Pref)
    assert np.max(np.abs(P[0].cpu().numpy() - Pref)) / np.max(np.abs(Pref)) < 5e-6
    assert abs(float(pmax[0]) - Pref.max()) / Pref.max() < 5e-6
    peaks, count = phy.peak_pick(P, pmax, r.shape[1], 16)
    assert np.array_equal(peaks[0, : int(count[0])].cpu().numpy(), g["peaks"])


@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_transmit(name, known_sequence):
    """rows 3-5: fused transmit chain against the reference's transmit() output (same RNG draws)."""
    torch = _torch()
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    phy = _phy(p)
    enc = np.concatenate([np.bitwise_xor(g["bits_in"].astype(np.int64),
                                         np.tile(known_sequence[: 2 * p.Nd], len(g["bits_in"]) // (2 * p.Nd) + 1)[: len(g["bits_in"])]),
                          g["pad"].astype(np.int64)])
    npk = len(enc) // phy.bits_per_packet
    packed = np.zeros((npk, phy.bits_stride), np.uint8)
    pb = np.packbits(enc.astype(np.uint8).reshape(npk, -1), axis=1)
    packed[:, : pb.shape[1]] = pb
    fill = torch.from_numpy(g["filler"].astype(np.complex64)).cuda().reshape(1, -1) if p.K > p.Nd else None
    out = phy.tx_modulate(torch.from_numpy(packed).cuda().reshape(1, npk, -1), fill, 1, npk)[0].cpu().numpy()
    assert out.shape == g["tx"].shape
    assert np.max(np.abs(out - g["tx"])) < 2e-7 * max(1.0, np.max(np.abs(g["tx"])) / 0.2)
    chirp = phy.sync_chirp().cpu().numpy()
    assert np.max(np.abs(chirp - orc.sync_chirp(p))) < 1e-7
    # encode("XOR") fused into the kernel (gf3_tx_encode_modulate): the un-encoded bits, padding pre-XORed with the known bits
    dbs = 2 * p.Nd
    nb = len(g["bits_in"])
    raw = np.concatenate([g["bits_in"].astype(np.int64),
                          np.bitwise_xor(g["pad"].astype(np.int64), known_sequence[:dbs][(nb + np.arange(len(g["pad"]))) % dbs])])
    packed2 = np.zeros((npk, phy.bits_stride), np.uint8)
    pb2 = np.packbits(raw.astype(np.uint8).reshape(npk, -1), axis=1)
    packed2[:, : pb2.shape[1]] = pb2
    out2 = phy.tx_modulate(torch.from_numpy(packed2).cuda().reshape(1, npk, -1), fill, 1, npk, xor=True)[0].cpu().numpy()
    assert np.array_equal(out2, out), "device-side encode differs from the host-encoded transmit"


# ----------------------------------------------------------------------------- known answers
def test_kat1_gr5ch1_dropin_receive(known_sequence, capsys):
    """The reference's one published known answer through the drop-in OFDM module:
    receiver("A2","XOR").receive(gr5ch1_signal.wav) -> the reference's 1 512 000 bits, hence BER
    0.023375665289067146 against gr5ch1.bmp (Final System Test.ipynb:85-169)."""
    _torch()
    import OFDM
    g = load_golden("kat1_gr5ch1.npz")
    rx = OFDM.receiver(mode="A2", encoding="XOR")
    details = {}
    bits, Hs0, He0 = rx.receive(g["wav_u8"] / 1.0, _details=details)
    printed = capsys.readouterr().out
    assert "Number of received OFDM symbols:    540" in printed and "Number of received bits:            1512000" in printed
    assert np.array_equal(details["peaks"], g["peaks"])
    np.testing.assert_allclose(details["slope"], g["slope"], rtol=0, atol=2e-7)
    hscale = np.max(np.abs(g["Hs"]))
    assert np.max(np.abs(details["Hs"] - g["Hs"])) / hscale < 2e-6
    assert np.max(np.abs(Hs0 - g["Hs"][0])) / hscale < 2e-6 and np.max(np.abs(He0 - g["He"][0])) / hscale < 2e-6
    dc = np.arange(100, 1500) - 1
    err = _rel_err(details["eq"][g["eq_rows"]][:, dc], g["eq_sel"][:, dc])
    assert err.max() < EQ_RTOL, err.max()
    ref_bits = np.unpackbits(g["bits_packed"])[:1512000]
    diff = np.flatnonzero(bits != ref_bits)
    # every differing decision must be one of the reference's own points within 1e-5 of a decision boundary
    # (the golden lists the points below 1e-4 with their margins: 2 of the 756 000 are below 1e-5)
    near5 = g["near_boundary"][g["near_margin"] < BOUNDARY_TOL]
    assert len(near5) == 2
    near = {(int(a), int(b)) for a, b in near5}
    for i in diff:
        assert ((i // 2) // 1400, (i // 2) % 1400) in near, "bit %d differs away from a decision boundary" % i
    with capsys.disabled():
        print("\nKAT-1: %d of 1512000 bits differ from the reference's (allowed only at its %d points within 1e-5 of a boundary)"
              % (len(diff), len(near5)))
    tx_bits = orc.load_file_bits("gr5ch1.bmp", g["bmp"])
    nerr = int(np.sum(tx_bits != bits[: len(tx_bits)]))
    assert abs(nerr - 24525) <= len(diff)
    if len(diff) == 0:
        assert repr(nerr / len(tx_bits)) == "0.023375665289067146"
        name, data = OFDM.save_file.__wrapped__(bits) if hasattr(OFDM.save_file, "__wrapped__") else _save(bits)
        assert name == "gr5ch1.bmp" and np.array_equal(data, g["file_payload"])


def _save(bits):
    name, size, payload = orc.save_file_bytes(bits)
    return name, payload


def test_kat3_weekend_known_channel(known_sequence):
    """Weekend Challenge.ipynb:162-310: N=1024, CP=32, 350 symbols through the known 30-tap channel
    decode back to y5tv9o.wav byte for byte."""
    torch = _torch()
    import gf3b200
    from scipy.signal import lfilter
    g = load_golden("kat3_weekend.npz")
    wav, h = g["y5tv9o_wav"], g["gr5channel"]
    p = orc.Params(N=1024, cp=32, lo=1, hi=512, known_sequence=known_sequence, encoding="None")
    bits = orc.load_file_bits("y5tv9o.wav", wav)
    nsym = 350
    bits = np.concatenate([bits, np.zeros(nsym * 2 * p.K - len(bits), dtype=np.uint8)])
    X = np.zeros((nsym, p.N), dtype=complex)
    X[:, 1:p.K + 1] = orc.qpsk_map(bits.reshape(nsym, p.K, 2))
    X[:, -np.arange(1, p.K + 1)] = np.conj(X[:, 1:p.K + 1])
    y = lfilter(h, 1.0, orc.add_cp(p, np.fft.ifft(X).real).reshape(-1)).astype(np.float32)
    phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=0, packet_len=nsym, known_sequence=known_sequence)
    Hinv = torch.from_numpy((1.0 / np.fft.fft(h, p.N)[1:p.K + 1]).astype(np.complex64)).cuda()
    packed, eq = phy.rx_known_channel(torch.from_numpy(y).cuda(), 1, Hinv, want_eq=True)
    got = phy.unpack_bits(packed)
    ref_bits, ref_eq = orc.known_channel_decode(p, y.astype(np.float64).reshape(nsym, -1), np.fft.fft(h, p.N))
    assert _rel_err(eq.cpu().numpy().reshape(nsym, -1), ref_eq).max() < EQ_RTOL
    assert np.array_equal(got, ref_bits) and np.array_equal(got, bits)
    name, size, data = orc.save_file_bytes(got)
    assert name == "y5tv9o.wav" and np.array_equal(data, wav)


def test_sync_quirk_wipeout(known_sequence):
    """OFDM.py:366-370: fewer than 2 samples after the final chirp wipes every detection."""
    torch = _torch()
    g = load_golden("sync_quirk.npz")
    p = oracle_params(g["cfg"], known_sequence, encoding="None")
    p.fit_lo, p.fit_hi = 10, 100
    phy = _phy(p)
    for trail in (0, 1, 2, 3):
        r = np.concatenate([np.zeros(100, np.float32), g["sig"], np.zeros(trail, np.float32)]).reshape(1, -1)
        P, pmax = phy.xcorr(torch.from_numpy(r).cuda())
        peaks, count = phy.peak_pick(P, pmax, r.shape[1], 8)
        assert np.array_equal(peaks[0, : int(count[0])].cpu().numpy(), g["peaks_trail%d" % trail]), trail


def test_peak_pick_paths_agree(known_sequence):
    """The one-kernel path (batches of at least two waves of streams) and the mark + scan path
    (validated against the goldens above) walk the same detections: many streams at different
    delays, noise levels above the threshold, truncated tails (wipe-out quirk), rows at 16-byte
    aligned and unaligned addresses."""
    torch = _torch()
    g = load_golden("sync_quirk.npz")
    p = oracle_params(g["cfg"], known_sequence, encoding="None")
    p.fit_lo, p.fit_hi = 10, 100
    phy = _phy(p)
    sig = torch.from_numpy(g["sig"].astype(np.float32)).cuda()
    B = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 37
    T = sig.numel() + 240
    gen = torch.Generator(device="cuda").manual_seed(99)
    r = 0.02 * torch.randn((B, T), device="cuda", generator=gen)
    for s in range(B):
        d = {20: 240, 21: 239}.get(s, (s * 7) % 200)                # 20, 21: no room after the last chirp (wipe-out)
        n = sig.numel() if s % 5 or s in (20, 21) else sig.numel() - (s % 11)   # some streams lose their tail
        r[s, d:d + n] += sig[:n] * (1.0 + 0.01 * (s % 13))
    if B > 10:
        r[10] = 0.3 * torch.randn(T, device="cuda", generator=gen)  # noise only: many candidates above 0.4 max
    P, pmax = phy.xcorr(r)
    one, one_n = phy.peak_pick(P, pmax, T, 8)
    parts = [phy.peak_pick(P[i:i + 64], pmax[i:i + 64], T, 8) for i in range(0, B, 64)]
    two = torch.cat([q[0] for q in parts]); two_n = torch.cat([q[1] for q in parts])
    assert torch.equal(one_n, two_n) and torch.equal(one, two)
    assert int((one_n > 0).sum()) > B // 2 and int(one_n.max()) >= 3 and int(one_n[20]) == 0
    for shift in (1, 2, 3):                                         # rows that are not 16-byte aligned
        buf = torch.zeros((B, P.stride(0) + 4), dtype=torch.float32, device="cuda")
        Pu = buf[:, shift:shift + P.shape[1]]
        Pu.copy_(P)
        o, n = phy.peak_pick(Pu, pmax, T, 8)
        assert torch.equal(n, one_n) and torch.equal(o, one), shift


# ----------------------------------------------------------------------------- properties at scale
@pytest.mark.parametrize("cfg", [
    dict(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, streams=256),      # BASELINE config C3 shape
    dict(N=4096, cp=704, lo=1, hi=2047, n_pilots=20, packet_len=180, streams=24),     # C4 / mode B1
    dict(N=4096, cp=224, lo=100, hi=1500, n_pilots=20, packet_len=180, streams=24),   # mode A2
    dict(N=256, cp=16, lo=3, hi=100, n_pilots=2, packet_len=33, streams=7),           # ragged: L % 16 != 0
])
def test_loopback_roundtrip_batch(cfg, known_sequence):
    """encode -> fused tx -> (ideal channel) -> sync -> fused rx recovers every bit, at the
    BASELINE.json shapes; also checks the batch path against per-stream calls."""
    torch = _torch()
    import gf3b200
    streams = cfg.pop("streams")
    phy = gf3b200.Phy(known_sequence=known_sequence, fit_lo=min(500, cfg["N"] // 8), fit_hi=min(1000, cfg["N"] // 4), **cfg)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    npk = 2
    bits = torch.randint(0, 256, (streams, npk, phy.bits_stride), dtype=torch.uint8, device="cuda", generator=gen)
    nbytes = phy.bits_per_packet // 8
    filler = None
    if phy.K > phy.Nd:
        f = torch.randint(0, 4, (streams, phy.K - phy.Nd), device="cuda", generator=gen)
        filler = (((1 - 2 * (f & 1)) + 1j * (1 - 2 * (f >> 1))) / np.sqrt(2)).to(torch.complex64)
    tx = phy.tx_modulate(bits, filler, streams, npk)
    r = torch.zeros((streams, tx.shape[1] + 10), dtype=torch.float32, device="cuda")
    r[:, 6:6 + tx.shape[1]] = tx
    P, pmax = phy.xcorr(r)
    peaks, count = phy.peak_pick(P, pmax, r.shape[1], 8)
    assert torch.all(count == npk + 1)
    expect = 6 + phy.chirp_len - 2 + torch.arange(npk + 1, device="cuda") * (phy.chirp_len + phy.pkt_samples)
    assert torch.all(peaks[:, : npk + 1] == expect[None, :])
    starts = (peaks[:, :npk] + 2) + (torch.arange(streams, device="cuda") * r.shape[1])[:, None]
    off = starts.reshape(-1).contiguous()
    Hs, He, slope = phy.rx_estimate(r.reshape(-1), streams * npk, off)
    out = phy.rx_demod(r.reshape(-1), streams * npk, Hs, He, slope, off, xor=False)
    assert torch.equal(out.reshape(streams, npk, -1)[:, :, :nbytes], bits[:, :, :nbytes])
    assert torch.all(out.reshape(streams, npk, -1)[:, :, (phy.bits_per_packet + 7) // 8:] == 0)      # pad bytes are zeroed
    # a single stream processed alone gives the same bytes (batching does not change results)
    o1 = phy.rx_demod(r[3].contiguous(), npk, Hs[3 * npk:3 * npk + npk], He[3 * npk:3 * npk + npk], slope[3 * npk:3 * npk + npk],
                      (peaks[3, :npk] + 2).contiguous(), xor=False)
    assert torch.equal(o1, out.reshape(streams, npk, -1)[3])


def test_channel_sim_and_ber_count(known_sequence):
    torch = _torch()
    import gf3b200
    from scipy.signal import lfilter
    phy = gf3b200.Phy(N=256, cp=16, lo=3, hi=100, n_pilots=2, packet_len=8, known_sequence=known_sequence, fit_lo=10, fit_hi=90)
    rng = np.random.default_rng(3)
    B, T, nt = 5, 10007, 30
    x = rng.normal(size=(B, T)).astype(np.float32)
    taps = rng.normal(size=(B, nt)).astype(np.float32)
    y = phy.channel_sim(torch.from_numpy(x).cuda(), torch.from_numpy(taps).cuda(), None, 1).cpu().numpy()
    ref = np.stack([lfilter(taps[b].astype(np.float64), 1.0, x[b].astype(np.float64)) for b in range(B)])
    assert np.max(np.abs(y - ref)) < 2e-5
    sigma = torch.full((B,), 0.5, device="cuda")
    z = torch.zeros((B, 400000), device="cuda")
    one = torch.ones((B, 1), device="cuda")
    n1 = phy.channel_sim(z, one, sigma, 11).cpu().numpy()
    n2 = phy.channel_sim(z, one, sigma, 11).cpu().numpy()
    n3 = phy.channel_sim(z, one, sigma, 12).cpu().numpy()
    assert np.array_equal(n1, n2) and not np.array_equal(n1, n3)
    assert abs(n1.std() - 0.5) < 2e-3 and abs(n1.mean()) < 2e-3
    assert abs(np.corrcoef(n1[0], n1[1])[0, 1]) < 0.01
    a = rng.integers(0, 256, 100001, dtype=np.uint8)
    b = rng.integers(0, 256, 100001, dtype=np.uint8)
    for nbits in (800008, 800003, 5):
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        phy.ber_count(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), nbits, cnt)
        ref_err = int(np.sum(np.unpackbits(a)[:nbits] != np.unpackbits(b)[:nbits]))
        assert cnt.cpu().tolist() == [ref_err, nbits]


def test_dropin_transmit_receive_file_roundtrip(known_sequence, tmp_path, monkeypatch):
    """Final System Test.ipynb flow through the drop-in: load_file -> transmit -> receive ->
    save_file recovers gr5ch1.bmp bit-exactly over an ideal channel (KAT-2), modes A2 and C2."""
    _torch()
    import OFDM
    g = load_golden("kat1_gr5ch1.npz")
    (tmp_path / "input_Files").mkdir()
    (tmp_path / "output_files").mkdir()
    g["bmp"].tofile(tmp_path / "input_Files" / "gr5ch1.bmp")
    monkeypatch.chdir(tmp_path)
    bits = OFDM.load_file("gr5ch1.bmp")
    assert len(bits) == 1049168                                            # Final System Test.ipynb:50
    for mode in ("A2", "C2"):
        np.random.seed(0)
        tx = OFDM.transmitter(mode=mode, encoding="XOR")
        sig = tx.transmit(bits)
        assert tx.no_packets == 3                                          # Final System Test.ipynb:52
        rx = OFDM.receiver(mode=mode, encoding="XOR")
        out, Hs, He = rx.receive(np.concatenate([np.zeros(1000), sig, np.zeros(500)]))
        assert len(out) == 1512000 and np.array_equal(out[: len(bits)], bits)
        name, data = OFDM.save_file(out)
        assert name == "gr5ch1.bmp" and np.array_equal(data, g["bmp"])
        assert np.array_equal(np.fromfile(tmp_path / "output_files" / "gr5ch1_received.bmp", dtype=np.uint8), g["bmp"])
    # recording cut inside the first packet: one detection, dropped as "terminating" -> the
    # reference's np.vstack([]) raises ValueError (OFDM.py:395,400); same here
    with pytest.raises(ValueError):
        rx.receive(np.concatenate([np.zeros(1000), sig[:600000]]))


def test_dropin_packed_file_path(known_sequence, tmp_path, monkeypatch):
    """file bytes -> transmit_bytes -> receive_bytes -> save_file_bytes with no bit array on the host gives
    the same waveform and the same bytes as the bit-array calls, and bit_errors() (device popcount) reproduces the
    notebook's BER of the real recording (Final System Test.ipynb:150-160: 24 525 errors, BER 0.0233756...)."""
    _torch()
    import OFDM
    g = load_golden("kat1_gr5ch1.npz")
    (tmp_path / "input_Files").mkdir()
    (tmp_path / "output_files").mkdir()
    g["bmp"].tofile(tmp_path / "input_Files" / "gr5ch1.bmp")
    monkeypatch.chdir(tmp_path)
    bits = OFDM.load_file("gr5ch1.bmp")
    fb = OFDM.load_file_bytes("gr5ch1.bmp")
    for mode, enc in (("A2", "XOR"), ("C2", "None")):
        tx = OFDM.transmitter(mode=mode, encoding=enc)
        np.random.seed(3)
        sig = tx.transmit(bits)
        np.random.seed(3)
        sig_b = tx.transmit_bytes(fb)
        assert np.array_equal(sig, sig_b)
        rx = OFDM.receiver(mode=mode, encoding=enc)
        r = np.concatenate([np.zeros(1000), sig, np.zeros(500)])
        out, Hs, He = rx.receive(r)
        out_b, Hs_b, He_b = rx.receive_bytes(r)
        assert out_b.dtype == np.uint8 and np.array_equal(out_b, np.packbits(out)) and np.array_equal(Hs, Hs_b)
        name, data = OFDM.save_file_bytes(out_b)
        assert name == "gr5ch1.bmp" and np.array_equal(data, g["bmp"])
        assert rx.bit_errors(out_b, fb) == (0, len(bits))
    # KAT-1: the real recording against the transmitted file, counted on the device
    rx = OFDM.receiver(mode="A2", encoding="XOR")
    got, _, _ = rx.receive_bytes(g["wav_u8"])
    assert np.array_equal(got, g["bits_packed"][: len(got)])
    errors, n = rx.bit_errors(got, fb)
    assert (errors, n) == (24525, 1049168) and errors / n == 0.023375665289067146


def test_ber_sweep_sharding_invariance(known_sequence):
    """configs[4] / SURVEY 8e: the stream-sharded tx -> channel -> sync -> rx sweep gives identical
    counters however the streams are split across ranks (here: 3 emulated ranks on one GPU)."""
    _torch()
    import gf3b200
    from gf3b200.sweep import make_gpu_count_fn, sweep
    phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=4, packet_len=12, known_sequence=known_sequence, fit_lo=125, fit_hi=250)
    count = make_gpu_count_fn(phy)
    snrs = [4.0, 25.0]
    full = sweep(count, 10, snrs, 0, 1, chunk=4)
    parts = sum(sweep(count, 10, snrs, r, 3, chunk=3) for r in range(3))
    assert np.array_equal(full, parts)
    assert full[0, 1] == full[1, 1] == 10 * phy.bits_per_packet
    assert full[1, 2] == 0                                  # every chirp found at 25 dB
    assert full[1, 0] < full[0, 0]                          # BER falls with SNR
    assert full[1, 0] / full[1, 1] < 0.02


def test_results_independent_of_launch_geometry(known_sequence):
    """The same packets give byte-identical bits whether they are demodulated in one large launch
    (every persistent CTA walks many re-seed blocks) or a few at a time (fewer blocks than CTAs),
    on a noisy channel with a clock offset; also in the exact-rotation mode (large slope)."""
    torch = _torch()
    import gf3b200
    from gf3b200 import synth
    phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, known_sequence=known_sequence,
                      fit_lo=125, fit_hi=250)
    b = synth.make_batch(phy, 1536, 1, snr_db=12.0, seed=5)
    sym = synth.packets_from_streams(phy, b)
    n = sym.shape[0]
    Hs, He, slope0 = phy.rx_estimate(sym.reshape(-1), n)
    for extra in (0.02, 0.9):                                  # fast (tan) rotation / exact rotation
        slope = slope0 + extra
        big = phy.rx_demod(sym.reshape(-1), n, Hs, He, slope)
        for lo_, hi_ in ((0, 3), (700, 764), (1500, 1536)):
            small = phy.rx_demod(sym[lo_:hi_].reshape(-1), hi_ - lo_, Hs[lo_:hi_].contiguous(), He[lo_:hi_].contiguous(),
                                 slope[lo_:hi_].contiguous())
            assert torch.equal(small, big[lo_:hi_]), (extra, lo_)


def test_pcm_ingest(known_sequence, capsys):
    """uint8 / int16 PCM ingest: exact value conversion on the device; the real recording decoded
    from its native uint8 samples gives the same bits as from `r/1.0` (Final System Test.ipynb:86)."""
    torch = _torch()
    import gf3b200
    import OFDM
    from gf3b200.host import HostReceiver
    phy = gf3b200.Phy(N=256, cp=16, lo=3, hi=100, n_pilots=2, packet_len=8, known_sequence=known_sequence, fit_lo=10, fit_hi=90)
    rng = np.random.default_rng(0)
    for dt, lo_, hi_ in ((np.uint8, 0, 256), (np.int16, -32768, 32768)):
        a = rng.integers(lo_, hi_, 100003).astype(dt)
        out = phy.pcm_to_f32(torch.from_numpy(a).cuda()).cpu().numpy()
        assert out.dtype == np.float32 and np.array_equal(out, a.astype(np.float32))
    g = load_golden("kat1_gr5ch1.npz")
    rx = OFDM.receiver(mode="A2", encoding="XOR")
    b1, _, _ = rx.receive(g["wav_u8"])               # native uint8
    b2, _, _ = rx.receive(g["wav_u8"] / 1.0)         # the notebook's float conversion
    capsys.readouterr()
    assert np.array_equal(b1, b2)
    # host-buffer receiver fed with int16 packets == float32 packets holding the same values
    gl = load_golden("stage_w1024.npz")
    phy2 = _phy(oracle_params(gl["cfg"], known_sequence))
    starts = (gl["peaks"] + 2)[:-1]
    pk = np.stack([gl["r_i16"][s:s + phy2.pkt_samples] for s in starts])
    h16 = torch.from_numpy(pk.astype(np.int16)).pin_memory()
    h32 = torch.from_numpy(pk.astype(np.float32)).pin_memory()
    o16 = HostReceiver(phy2, len(starts), sample_dtype=torch.int16).run(h16).clone()
    o32 = HostReceiver(phy2, len(starts)).run(h32).clone()
    torch.cuda.synchronize()
    assert torch.equal(o16, o32)
    assert np.array_equal(phy2.unpack_bits(o16), gl["bits"])


def test_fused_receive_is_deterministic_and_equals_two_launches(known_sequence):
    """A large batch through the one-launch chain twice, and through gf3_rx_estimate + gf3_rx_demod:
    byte-identical packed bits (persistent CTAs split packets between them, estimate some packets
    twice and skip barriers around the flush -- any race would show up as a difference)."""
    torch = _torch()
    import gf3b200
    from gf3b200 import synth
    phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, known_sequence=known_sequence,
                      fit_lo=125, fit_hi=250)
    b = synth.make_batch(phy, 1200, 1, snr_db=14.0, seed=77)
    sym = synth.packets_from_streams(phy, b).contiguous()
    n = sym.shape[0]
    flat = sym.reshape(-1)
    one, Hs1, He1, sl1 = phy.rx_receive(flat, n, xor=True)
    again, _, _, sl2 = phy.rx_receive(flat, n, xor=True)
    Hs, He, slope = phy.rx_estimate(flat, n)
    two = phy.rx_demod(flat, n, Hs, He, slope, xor=True)
    torch.cuda.synchronize()
    assert torch.equal(one, again) and torch.equal(sl1, sl2)
    assert float((Hs1 - Hs).abs().max() / Hs.abs().max()) < 1e-6           # the fused estimate sums the pilots in another order
    diff = (one != two).any(dim=1)
    # decisions may differ only where the two estimates' rounding moves a point across a boundary
    assert int(diff.sum()) <= max(1, n // 100), int(diff.sum())
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    phy.ber_count(one.contiguous(), two.contiguous(), n * phy.bits_stride * 8, cnt)
    assert int(cnt[0]) <= n // 50, int(cnt[0])


def test_schmidlcox_prefix_sum_kernel(known_sequence, capsys):
    """OFDM.py:376-387 on the device (gf3_schmidlcox): the reference's indices on the golden signals (float32, and int16
    read natively), a batch against the oracle, and the drop-in method; a recording that is too short raises IndexError."""
    torch = _torch()
    import gf3b200
    import OFDM
    from oracle.make_golden import sc_signal
    g = load_golden("sync_schmidlcox.npz")
    p = orc.Params.from_mode("A2", known_sequence=known_sequence)
    rx = OFDM.receiver(mode="A2", encoding="XOR")
    for seed in (17, 18):
        r = sc_signal(seed)
        assert rx.schmidlcox_method(r) == int(g["index_seed%d" % seed])
        assert rx.schmidlcox_method(np.round(r * 8000).astype(np.int16)) == int(g["index_i16_seed%d" % seed])
    phy = gf3b200.Phy(known_sequence=known_sequence)
    rng = np.random.default_rng(5)
    batch = np.stack([sc_signal(100 + i) * (1 + 0.1 * i) for i in range(6)])
    batch[3] = rng.normal(0, 1.0, batch.shape[1])                       # noise only: argmax anywhere
    q = np.round(batch * 4000).astype(np.int16)
    idx, val = phy.schmidlcox(torch.from_numpy(q).cuda(), 5 * 48000)
    ref = [orc.schmidlcox_method(p, q[i].astype(np.float64)) - p.N + 1 for i in range(6)]
    assert idx.cpu().tolist() == ref
    with pytest.raises(IndexError):
        rx.schmidlcox_method(np.zeros(5 * 48000))
