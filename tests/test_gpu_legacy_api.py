"""The OLD API of the reference's notebooks (older OFDM.py revisions whose code is not in the repository):
CamG(N, cp, "QPSK"), module-level FFT / IFFT / equalise(Y, H), keyword receiver(ofdm_symbol_size=, ...).  The
Weekend-Challenge notebook's own code cells (tests/golden/notebook_cells.json, extracted from the reference by
oracle/make_golden.py cells) run UNMODIFIED against the drop-in on the regenerated KAT-3 input and must write
y5tv9o.wav byte for byte."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import gf3_oracle as orc

pytestmark = pytest.mark.gpu


class _Plot:
    """matplotlib stand-in for the notebook's plotting lines (matplotlib is not installed on the box)."""

    def __getattr__(self, name):
        return lambda *a, **k: None


def _cells(nb):
    with open(os.path.join(GOLDEN, "notebook_cells.json")) as f:
        return json.load(f)[nb]


def _run(cells, ns, skip=()):
    for c in cells:
        if c["index"] in skip:
            continue
        src = "\n".join(line for line in c["source"].split("\n") if not line.lstrip().startswith("%"))   # IPython magics
        exec(compile(src, "cell %d" % c["index"], "exec"), ns)


def test_weekend_challenge_notebook_cells_unmodified(known_sequence, tmp_path, monkeypatch):
    import torch
    assert torch.cuda.is_available()
    from scipy.signal import lfilter
    g = load_golden("kat3_weekend.npz")
    wav, h = g["y5tv9o_wav"], g["gr5channel"]
    # regenerate the missing handouts/gr5file.csv: the file the notebook decodes
    p = orc.Params(N=1024, cp=32, lo=1, hi=512, known_sequence=known_sequence, encoding="None")
    payload = wav[44:]                                             # 44-byte wav header + 44 612 data bytes
    assert len(payload) == 44612
    bits = orc.load_file_bits("y5tv9o.wav", payload)
    nsym = 350
    bits = np.concatenate([bits, np.zeros(nsym * 2 * p.K - len(bits), dtype=np.uint8)])
    X = np.zeros((nsym, p.N), dtype=complex)
    X[:, 1:p.K + 1] = orc.qpsk_map(bits.reshape(nsym, p.K, 2))
    X[:, -np.arange(1, p.K + 1)] = np.conj(X[:, 1:p.K + 1])
    y = lfilter(h, 1.0, orc.add_cp(p, np.fft.ifft(X).real).reshape(-1))
    (tmp_path / "handouts").mkdir()
    (tmp_path / "sound_files").mkdir()
    with open(tmp_path / "handouts" / "gr5file.csv", "w") as f:
        f.write(" ".join(repr(float(v)) for v in y) + "\n")
    with open(tmp_path / "handouts" / "gr5channel.csv", "w") as f:
        f.write(" ".join(repr(float(v)) for v in h) + "\n")
    monkeypatch.chdir(tmp_path)
    ns = {}
    exec("from OFDM import *", ns)                                 # cell 1 without its %matplotlib magic
    ns["plt"] = _Plot()
    import types
    ipd = types.ModuleType("IPython.display")
    ipd.Audio = lambda *a, **k: None
    monkeypatch.setitem(__import__("sys").modules, "IPython", types.ModuleType("IPython"))
    monkeypatch.setitem(__import__("sys").modules, "IPython.display", ipd)
    cells = _cells("Weekend Challenge.ipynb")
    _run(cells, ns, skip=(1,))
    assert ns["wc"].K == 1024 and ns["symbols"].shape == (350, 511) and ns["data_P"].shape == (350, 511, 2)
    assert ns["file_name"] == "y5tv9o.wav" and ns["file_size"] == "44612"
    assert np.array_equal(ns["file_data"], payload)
    written = np.fromfile(tmp_path / "sound_files" / "y5tv9o.wav", dtype=np.uint8)
    assert np.array_equal(written, wav), "the notebook's own wavfile.write of the decoded payload reproduces sound_files/y5tv9o.wav"


def test_old_api_fft_ifft_equalise_roundtrip(known_sequence):
    """Initial OFDM Test.ipynb's single-symbol loop-back (cells 3-25; the saved run stops at an IndexError in cell 13, the
    flow itself is bits -> SP -> map -> OFDM_symbol -> IFFT -> add_cp -> remove_cp -> FFT -> equalise -> get_data -> demap
    -> PS == bits), on the device-backed shims, N = 64 / CP = 16 as in the notebook and N = 1024."""
    import OFDM
    for N, cp in ((64, 16), (1024, 32)):
        test = OFDM.CamG(N, cp, "QPSK")
        rng = np.random.default_rng(N)
        bits = rng.integers(0, 2, test.bits_per_symbol)
        QPSK = test.map(test.SP(bits))
        OFDM_data = test.OFDM_symbol(QPSK)
        assert len(OFDM_data) == N
        OFDM_time = OFDM.IFFT(OFDM_data)
        assert np.max(np.abs(OFDM_time - np.fft.ifft(OFDM_data))) < 2e-7
        OFDM_rx = test.remove_cp(test.add_cp(OFDM_time))
        OFDM_demod = OFDM.FFT(OFDM_rx)
        assert np.max(np.abs(OFDM_demod - np.fft.fft(OFDM_rx))) < 1e-5
        OFDM_demod = OFDM.equalise(OFDM_demod, np.ones(len(test.all_carriers)))
        bits_PS, decisions = test.demap(test.get_data(OFDM_demod))
        np.testing.assert_array_equal(test.PS(bits_PS), bits)                 # the notebook's own assert (cell 25)
    x = rng.standard_normal((5, 256))
    assert np.max(np.abs(OFDM.FFT(x) - np.fft.fft(x))) < 2e-5
    with pytest.raises(ValueError):
        OFDM.FFT(x + 1j)


def test_old_api_keyword_transmitter_receiver_loopback(known_sequence, capsys):
    """Audio.ipynb:60-161 without the sound card: transmitter(N, cp, "QPSK") / receiver(N, cp, "QPSK", pilot_sequence=)
    with the caller's own known bits, transmit -> receive over an ideal channel, BER computed as in the notebook."""
    import OFDM
    ofdm_symbol_size, cp_length, modulation = 1024, 128, "QPSK"
    np.random.seed(3)
    tx = OFDM.transmitter(ofdm_symbol_size, cp_length, modulation)
    no_bits = tx.bits_per_symbol * 50
    bits = np.random.binomial(n=1, p=0.5, size=(no_bits,))
    known_bits = np.random.binomial(n=1, p=0.5, size=(tx.bits_per_symbol,))
    tx.pilot_sequence = known_bits
    signal = tx.transmit(bits, graph_output=False)
    rx = OFDM.receiver(ofdm_symbol_size, cp_length, modulation, pilot_sequence=known_bits)
    rx_bits = rx.receive(np.concatenate([np.zeros(777), signal, np.zeros(100)]))
    capsys.readouterr()
    errs = np.sum(abs(bits - rx_bits[:len(bits)]))
    assert errs / len(bits) == 0.0
    assert rx.Hest is not None and len(rx.Hest) == rx.K
    kw = OFDM.receiver(ofdm_symbol_size=4096, cp_length=0, modulation="QPSK", fs=48000, end_sync=False)
    assert (kw.ofdm_symbol_size, kw.cp_length, kw.K, kw.data_carriers_per_symbol, kw.end_sync) == (4096, 0, 2047, 2047, False)
    sc = OFDM.receiver(ofdm_symbol_size=4096, cp_length=100, modulation="QPSK", fs=48000, no_pilots=100,
                       pilot_sequence=np.zeros(4094, dtype=int), sync_method='schmidlcox', end_sync=True)
    with pytest.raises(NotImplementedError):
        sc.receive(np.zeros(10000))
