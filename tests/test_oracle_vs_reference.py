"""Live comparison of the numpy oracle with the UNMODIFIED reference (/root/reference/OFDM.py), stage by
stage, on small seeded cases.  Runs only where the reference tree exists (the build container);
skipped on the GPU box -- there the committed goldens (tests/golden, made by oracle/make_golden.py
from these same reference calls) carry the pinning."""
import numpy as np
import pytest

from oracle import gf3_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

CASES = [
    # N, cp, lo, hi, P, L, packets, seed
    (1024, 32, 1, 512, 3, 6, 2, 21),
    (2048, 64, 40, 700, 2, 5, 1, 22),
]


@pytest.mark.parametrize("N,cp,lo,hi,P,L,npk,seed", CASES)
def test_oracle_equals_reference_live(N, cp, lo, hi, P, L, npk, seed, known_sequence, tmp_path, monkeypatch):
    import os
    cwd = os.getcwd()
    try:
        kw = dict(no_pilots=P, packet_length=L, N=N, cp=cp, lo=lo, hi=hi)
        tx = ref_shim.make("transmitter", "A1", "XOR", **kw)
        rx = ref_shim.make("receiver", "A1", "XOR", **kw)
        p = orc.Params(N=N, cp=cp, lo=lo, hi=hi, n_pilots=P, packet_len=L, known_sequence=known_sequence, encoding="XOR")
        rng = np.random.default_rng(seed)
        bits = rng.integers(0, 2, 2 * (hi - lo) * L * npk - 11)
        np.random.seed(seed)
        with ref_shim.quiet():
            sig = tx.transmit(bits)                                   # OFDM.py:296-343
        np.random.seed(seed)
        sig_o = orc.transmit(p, bits)
        assert np.max(np.abs(sig - sig_o)) < 1e-15
        assert np.max(np.abs(tx.sync_chirp() - orc.sync_chirp(p))) < 1e-12
        h = np.array([1.0, 0.4, -0.2, 0.1])
        r = np.convolve(sig, h)[: len(sig)]
        r = np.concatenate([np.zeros(77), r, np.zeros(9)]) + rng.normal(0, 2e-3, len(sig) + 86)
        with ref_shim.quiet():
            zeros = rx.chirp_method(r)                                # OFDM.py:356-372
            rx_cp = rx.get_symbols(r, zeros)                          # OFDM.py:391-403
            data, sp, ep = rx.get_data(np.fft.fft(rx.remove_cp(rx_cp)))
            eq, Hs, He, Hest = rx.equalise(data, sp, ep)              # OFDM.py:422-480
            par, hard = rx.demap(eq[:, rx.data_carriers - 1])         # OFDM.py:484-500
            out_bits = rx.decode(rx.PS(par))                          # OFDM.py:504-505, 541-544
        o = orc.receive(p, r, want_eq=True)
        assert np.array_equal(np.where(zeros)[0], o["peaks"])
        assert np.allclose(o["Hs"], Hs, rtol=1e-12, atol=0) and np.allclose(o["He"], He, rtol=1e-12, atol=0)
        assert np.allclose(o["eq"], eq, rtol=1e-10, atol=1e-13)
        assert np.array_equal(o["bits"], out_bits)
        assert np.mean(out_bits[: len(bits)] != bits) < 0.2      # (the reference's own decode errs at the band edge here)
    finally:
        os.chdir(cwd)
