"""Pin the numpy oracle (oracle/gf3_oracle.py) against golden vectors produced by the unmodified
reference (oracle/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import STAGE_NAMES, kat4_regenerate, load_golden, oracle_params
from oracle import gf3_oracle as orc


def test_known_sequence_matches_reference(known_sequence):
    g = load_golden("kat1_gr5ch1.npz")
    assert np.array_equal(known_sequence, g["known_sequence"])
    assert len(known_sequence) == 4096


def test_kat1_gr5ch1_full_receive(known_sequence):
    """Final System Test.ipynb:85-169: 540 symbols, 1 512 000 bits, BER 0.023375665289067146."""
    g = load_golden("kat1_gr5ch1.npz")
    p = orc.Params.from_mode("A2", known_sequence=known_sequence)
    r = g["wav_u8"] / 1.0
    out = orc.receive(p, r, want_eq=True)
    assert np.array_equal(out["peaks"], g["peaks"])
    assert np.array_equal(out["peaks"], [71011, 1042993, 2014975, 2986958])
    assert len(out["bits"]) == 1512000
    np.testing.assert_allclose(out["slope"], g["slope"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["Hs"], g["Hs"], rtol=1e-13)
    np.testing.assert_allclose(out["He"], g["He"], rtol=1e-13)
    np.testing.assert_allclose(out["eq"][g["eq_rows"]], g["eq_sel"], rtol=1e-11)
    packed = np.packbits(out["bits"])
    assert hashlib.sha256(packed.tobytes()).hexdigest() == str(g["bits_sha256"])
    assert np.array_equal(packed, g["bits_packed"])
    tx_bits = orc.load_file_bits("gr5ch1.bmp", g["bmp"])
    nerr = int(np.sum(tx_bits != out["bits"][: len(tx_bits)]))
    assert nerr == 24525 == int(g["n_bit_errors"])
    assert repr(nerr / len(tx_bits)) == "0.023375665289067146"
    name, size, payload = orc.save_file_bytes(out["bits"])
    assert (name, size) == ("gr5ch1.bmp", "131128")
    assert np.array_equal(payload, g["file_payload"])


@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_receive(name, known_sequence):
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    r = g["r_i16"].astype(np.float64)
    out = orc.receive(p, r, want_eq=True)
    assert np.array_equal(out["peaks"], g["peaks"])
    np.testing.assert_allclose(out["slope"], g["slope"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(out["Hs"], g["Hs"], rtol=1e-12)
    np.testing.assert_allclose(out["He"], g["He"], rtol=1e-12)
    np.testing.assert_allclose(out["eq"], g["eq"], rtol=1e-10)
    assert np.array_equal(out["bits"], g["bits"])
    P, L = p.n_pilots, p.packet_len
    rx_cp, _ = orc.get_symbols(p, r, orc.chirp_method(p, r))
    ofdm = orc.rx_fft(p, rx_cp)
    np.testing.assert_allclose(ofdm[:, [0, P, P + L - 1, 2 * P + L - 1], 1:p.N // 2], g["ofdm_sel"], rtol=1e-12)


@pytest.mark.parametrize("name", STAGE_NAMES)
def test_stage_transmit(name, known_sequence):
    """transmit() with the reference's RNG draw order (binomial padding, then choice filler)."""
    g = load_golden("stage_%s.npz" % name)
    p = oracle_params(g["cfg"], known_sequence)
    np.random.seed(int(g["seed"]))
    tx = orc.transmit(p, g["bits_in"].astype(np.int64))
    assert tx.shape == g["tx"].shape
    np.testing.assert_allclose(tx[:4096], g["tx_f64_head"], rtol=0, atol=1e-16)
    np.testing.assert_allclose(tx, g["tx"].astype(np.float64), rtol=0, atol=2e-8)   # golden stored as f32
    # the pieces the GPU path takes from the host RNG
    np.random.seed(int(g["seed"]))
    enc = orc.encode(p, g["bits_in"].astype(np.int64))
    filler = orc.random_qpsk(p)
    assert np.array_equal(enc[len(g["bits_in"]):], g["pad"])
    assert np.array_equal(filler, g["filler"])


def test_sync_chirp_closed_form(known_sequence):
    from scipy.signal import chirp
    for mode in ("A2", "B1", "C3"):
        p = orc.Params.from_mode(mode, known_sequence=known_sequence)
        t = np.linspace(0, p.chirp_length / p.fs, p.chirp_length)
        ref = chirp(t, f0=p.f0, f1=p.f1, t1=p.chirp_length / p.fs, method="linear") / 5
        np.testing.assert_allclose(orc.sync_chirp(p), ref, rtol=0, atol=1e-12)


def test_sync_quirk_wipeout(known_sequence):
    """OFDM.py:366-370: < 2 trailing samples after the last chirp wipes every detection."""
    g = load_golden("sync_quirk.npz")
    p = oracle_params(g["cfg"], known_sequence, encoding="None")
    sig = g["sig"].astype(np.float64)
    for trail in (0, 1, 2, 3):
        r = np.concatenate([np.zeros(100), sig, np.zeros(trail)])
        peaks = np.where(orc.chirp_method(p, r))[0]
        assert np.array_equal(peaks, g["peaks_trail%d" % trail]), trail
    assert len(g["peaks_trail1"]) == 0 and len(g["peaks_trail2"]) == 3


def test_demap_equals_min_distance():
    rng = np.random.default_rng(0)
    s = rng.normal(size=(50, 64)) + 1j * rng.normal(size=(50, 64))
    s[0, :8] = [0, 1, -1, 1j, -1j, 1 + 0j, 0 + 0j, -0.0 - 0.0j]        # ties -> first minimum
    assert np.array_equal(orc.demap(s), orc.demap_min_distance(s))


def test_kat3_weekend_known_channel(known_sequence):
    """Weekend Challenge.ipynb:162-310: N=1024, CP=32, 350 symbols, known 30-tap channel ->
    y5tv9o.wav (44 612 payload bytes)."""
    from scipy.signal import lfilter
    g = load_golden("kat3_weekend.npz")
    wav, h = g["y5tv9o_wav"], g["gr5channel"]
    p = orc.Params(N=1024, cp=32, lo=1, hi=512, known_sequence=known_sequence, encoding="None")
    payload = wav[44:]                                    # the notebook's output is header-less data
    bits = orc.load_file_bits("y5tv9o.wav", wav)
    nsym = int(np.ceil(len(bits) / (2 * p.K)))
    assert nsym == 350
    bits = np.concatenate([bits, np.zeros(nsym * 2 * p.K - len(bits), dtype=np.uint8)])
    X = np.zeros((nsym, p.N), dtype=complex)
    X[:, 1:p.K + 1] = orc.qpsk_map(bits.reshape(nsym, p.K, 2))
    X[:, -np.arange(1, p.K + 1)] = np.conj(X[:, 1:p.K + 1])
    x = orc.add_cp(p, np.fft.ifft(X).real)
    y = lfilter(h, 1.0, x.reshape(-1)).reshape(nsym, p.sym_len)
    out_bits, eq = orc.known_channel_decode(p, y, np.fft.fft(h, p.N))
    assert np.array_equal(out_bits, bits)
    name, size, data = orc.save_file_bytes(out_bits)
    assert name == "y5tv9o.wav" and int(size) == len(wav)
    assert np.array_equal(data, wav) and len(payload) == 44612


def test_kat4_gr5ch2_long_recording(known_sequence):
    """KAT-4 = BASELINE.json configs[1]: the reference transmits input_Files/gr5ch2.wav in mode
    A2 (29 packets, 28.2 M samples), Handouts/gr5channel.csv FIR + AWGN + int16; the reference's own
    receive() output is the golden (OFDM.py:296-343, 581-657).  The oracle must regenerate the recording
    from the seed and decode it to the very same bits, sync indices, slopes and constellation."""
    g = load_golden("kat4_gr5ch2.npz")
    p, bits_in, r = kat4_regenerate(g, known_sequence)
    assert len(r) == int(g["n_samples"]) == 28214698
    out = orc.receive(p, r.astype(np.float64), want_eq=True)
    assert np.array_equal(out["peaks"], g["peaks"]) and len(out["peaks"]) == 30
    np.testing.assert_allclose(out["slope"], g["slope"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["Hs"][0], g["Hs0"], rtol=1e-12)
    np.testing.assert_allclose(out["He"][28], g["He28"], rtol=1e-12)
    np.testing.assert_allclose(out["eq"][g["eq_rows"]], g["eq_sel"], rtol=1e-10)
    assert len(out["bits"]) == int(g["n_bits"]) == 29 * 180 * 2800
    assert hashlib.sha256(np.packbits(out["bits"]).tobytes()).hexdigest() == str(g["bits_sha256"])
    assert int(np.sum(out["bits"][: len(bits_in)] != bits_in)) == int(g["n_bit_errors"]) == 111558
    name, size, _ = orc.save_file_bytes(out["bits"])
    assert name == str(g["file_name"]) == "gr5ch2.wav" and int(size) == len(g["payload"])


def test_schmidlcox_metric(known_sequence):
    """receiver.schmidlcox_method (OFDM.py:376-387): the oracle's cumulative sum gives the reference's indices."""
    from oracle.make_golden import sc_signal
    g = load_golden("sync_schmidlcox.npz")
    p = orc.Params.from_mode("A2", known_sequence=known_sequence)
    for seed in (17, 18):
        r = sc_signal(seed)
        assert orc.schmidlcox_method(p, r) == int(g["index_seed%d" % seed])
        assert orc.schmidlcox_method(p, np.round(r * 8000).astype(np.int16).astype(np.float64)) == int(g["index_i16_seed%d" % seed])
    assert abs(int(g["index_seed17"]) - (91234 + 4096 - 1)) < 64          # locks on the repeated half-symbol
