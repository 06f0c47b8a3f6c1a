"""GPU edge cases: every supported symbol size, odd / zero cyclic prefix, single data carrier,
one-symbol packets, one pilot, empty batches, arbitrary packet offsets; every output buffer is
surrounded by canaries to catch out-of-bounds writes (compute-sanitizer is not available on the pool)."""
import ctypes

import numpy as np
import pytest

from conftest import assert_bits_match
from oracle import gf3_oracle as orc

pytestmark = pytest.mark.gpu
CANARY = 0x5A


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


class Guarded:
    """A device buffer with canary bytes on both sides of the region handed to the library."""

    def __init__(self, torch, nbytes, dtype, shape, pad=4096):
        self.torch, self.pad, self.nbytes = torch, pad, nbytes
        self.raw = torch.full((nbytes + 2 * pad,), CANARY, dtype=torch.uint8, device="cuda")
        self.view = self.raw[pad:pad + nbytes].view(dtype).reshape(shape)

    def check(self, what):
        t = self.torch
        assert bool(t.all(self.raw[: self.pad] == CANARY)), "%s: write before the buffer" % what
        assert bool(t.all(self.raw[self.pad + self.nbytes:] == CANARY)), "%s: write past the buffer" % what


CASES = [
    # N, cp, lo, hi, P, L, fit_lo, fit_hi, packets
    (64, 16, 1, 32, 2, 5, 4, 28, 9),
    (128, 0, 1, 64, 1, 16, 8, 56, 5),          # no cyclic prefix, one pilot
    (256, 7, 3, 100, 2, 33, 10, 90, 4),        # odd CP: every other symbol is 4-byte aligned only
    (512, 64, 200, 201, 3, 17, 20, 200, 3),    # a single data carrier
    (1024, 32, 1, 512, 4, 1, 125, 250, 6),     # one data symbol per packet
    (1024, 32, 17, 400, 20, 180, 125, 250, 2),
    (1024, 33, 1, 512, 3, 19, 125, 250, 3),    # odd symbol length: the two symbols of a warp differ in alignment
    (2048, 63, 1, 1024, 2, 11, 250, 500, 2),
    (2048, 64, 5, 900, 3, 65, 250, 500, 2),    # crosses the 64-symbol re-seed period
    (4096, 224, 100, 1500, 4, 21, 500, 1000, 2),
    (4096, 1184, 1, 2047, 2, 9, 500, 1000, 1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%d_cp%d_%d-%d_P%d_L%d" % c[:6])
def test_receive_chain_edges_vs_oracle(case, known_sequence):
    torch = _torch()
    import gf3b200
    N, cp, lo, hi, P, L, flo, fhi, npk = case
    p = orc.Params(N=N, cp=cp, lo=lo, hi=hi, n_pilots=P, packet_len=L, known_sequence=known_sequence, fit_lo=flo, fit_hi=fhi)
    phy = gf3b200.Phy(N=N, cp=cp, lo=lo, hi=hi, n_pilots=P, packet_len=L, known_sequence=known_sequence, fit_lo=flo, fit_hi=fhi)
    rng = np.random.default_rng(N + cp + L)
    bits = rng.integers(0, 2, npk * p.data_bits_per_symbol * L)
    np.random.seed(N + L)
    tx = orc.transmit(p, bits)
    h = np.array([1.0, 0.3, -0.1, 0.05])[: max(1, min(4, cp + 1))]
    y = np.convolve(tx, h)[: len(tx)]
    y = np.concatenate([np.zeros(37), y, np.zeros(11)])
    y = (y + rng.normal(0, 2e-3, len(y))).astype(np.float32)
    ref = orc.receive(p, y.astype(np.float64), want_eq=True)
    assert len(ref["starts"]) == npk
    d = torch.from_numpy(y).cuda()
    off = torch.from_numpy(ref["starts"].astype(np.int64)).cuda()
    K = p.K
    # guarded outputs, raw C-ABI calls
    gHs = Guarded(torch, npk * K * 8, torch.complex64, (npk, K))
    gHe = Guarded(torch, npk * K * 8, torch.complex64, (npk, K))
    gsl = Guarded(torch, npk * 8, torch.float64, (npk,))
    gb = Guarded(torch, npk * phy.bits_stride, torch.uint8, (npk, phy.bits_stride))
    geq = Guarded(torch, npk * L * K * 8, torch.complex64, (npk, L, K))
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    from gf3b200._lib import check
    check(phy.lib.gf3_rx_estimate(phy._plan, vp(d), vp(off), npk, vp(phy.known), vp(gHs.view), vp(gHe.view), vp(gsl.view), st))
    check(phy.lib.gf3_rx_demod(phy._plan, vp(d), vp(off), npk, vp(gHs.view), vp(gHe.view), vp(gsl.view), vp(phy.xor2),
                               vp(gb.view), phy.bits_stride, vp(geq.view), st))
    torch.cuda.synchronize()
    for g, name in ((gHs, "Hs"), (gHe, "He"), (gsl, "slope"), (gb, "bits"), (geq, "eq")):
        g.check(name)
    hs = np.max(np.abs(ref["Hs"]))
    assert np.max(np.abs(gHs.view.cpu().numpy() - ref["Hs"])) / hs < 3e-6
    assert np.max(np.abs(gHe.view.cpu().numpy() - ref["He"])) / hs < 3e-6
    np.testing.assert_allclose(gsl.view.cpu().numpy(), ref["slope"], rtol=0, atol=5e-7)
    dc = p.data_carriers - 1
    eq = geq.view.cpu().numpy().reshape(-1, K)
    rel = np.abs(eq[:, dc] - ref["eq"][:, dc]) / np.maximum(np.abs(ref["eq"][:, dc]), 1e-30)
    assert rel.max() < 2e-4, rel.max()
    got = phy.unpack_bits(gb.view)
    pts = ref["eq"][:, dc].reshape(-1)
    nbad = assert_bits_match(got, ref["bits"], pts, "edge case")["n_diff"]      # only decisions within 1e-5 of a boundary may differ
    assert np.array_equal(got[: len(bits)], bits) or nbad > 0 or ref["bits"][: len(bits)].tolist() != bits.tolist()
    # pad bytes of every row are zero
    nb = (phy.bits_per_packet + 7) // 8
    assert bool(torch.all(gb.view[:, nb:] == 0))
    if phy.bits_per_packet % 8:
        mask = (1 << (8 - phy.bits_per_packet % 8)) - 1
        assert bool(torch.all((gb.view[:, nb - 1] & mask) == 0))
    # ---- the same chain as ONE launch (gf3_rx_receive), into fresh guarded buffers
    fHs = Guarded(torch, npk * K * 8, torch.complex64, (npk, K))
    fHe = Guarded(torch, npk * K * 8, torch.complex64, (npk, K))
    fsl = Guarded(torch, npk * 8, torch.float64, (npk,))
    fb = Guarded(torch, npk * phy.bits_stride, torch.uint8, (npk, phy.bits_stride))
    feq = Guarded(torch, npk * L * K * 8, torch.complex64, (npk, L, K))
    check(phy.lib.gf3_rx_receive(phy._plan, vp(d), vp(off), npk, vp(phy.known), vp(fHs.view), vp(fHe.view), vp(fsl.view),
                                 vp(phy.xor2), vp(fb.view), phy.bits_stride, vp(feq.view), st))
    torch.cuda.synchronize()
    for g, name in ((fHs, "fused Hs"), (fHe, "fused He"), (fsl, "fused slope"), (fb, "fused bits"), (feq, "fused eq")):
        g.check(name)
    assert np.max(np.abs(fHs.view.cpu().numpy() - ref["Hs"])) / hs < 3e-6
    assert np.max(np.abs(fHe.view.cpu().numpy() - ref["He"])) / hs < 3e-6
    np.testing.assert_allclose(fsl.view.cpu().numpy(), ref["slope"], rtol=0, atol=5e-7)
    relf = np.abs(feq.view.cpu().numpy().reshape(-1, K)[:, dc] - ref["eq"][:, dc]) / np.maximum(np.abs(ref["eq"][:, dc]), 1e-30)
    assert relf.max() < 2e-4, relf.max()
    gotf = phy.unpack_bits(fb.view)
    assert_bits_match(gotf, ref["bits"], pts, "edge case, fused")
    assert bool(torch.all(fb.view[:, nb:] == 0))


def test_empty_batches_and_bad_arguments(known_sequence):
    torch = _torch()
    import gf3b200
    from gf3b200 import _lib
    phy = gf3b200.Phy(N=256, cp=16, lo=3, hi=100, n_pilots=2, packet_len=8, known_sequence=known_sequence, fit_lo=10, fit_hi=90)
    d = torch.zeros(8, device="cuda")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    # zero packets / streams: success, nothing launched
    n0 = gf3b200.launch_count()
    assert phy.lib.gf3_rx_estimate(phy._plan, vp(d), None, 0, vp(phy.known), vp(d), vp(d), vp(d), None) == 0
    assert phy.lib.gf3_rx_demod(phy._plan, vp(d), None, 0, vp(d), vp(d), vp(d), None, vp(d), 16, None, None) == 0
    assert phy.lib.gf3_xcorr(phy._plan, vp(d), 8, 0, 8, vp(d), 2000, vp(d), vp(d), None) == 0
    assert phy.lib.gf3_tx_modulate(phy._plan, vp(d), phy.bits_stride, vp(d), vp(phy.known), 0, 1, vp(d), 10 ** 6, None) == 0
    assert phy.lib.gf3_rx_receive(phy._plan, vp(d), None, 0, vp(phy.known), vp(d), vp(d), vp(d), None, vp(d), 16, None, None) == 0
    assert phy.lib.gf3_eq_estimate(phy._plan, vp(d), vp(d), 0, vp(phy.known), vp(d), vp(d), vp(d), None) == 0
    assert phy.lib.gf3_eq_apply(phy._plan, vp(d), 0, vp(d), vp(d), vp(d), vp(d), None, None) == 0
    assert phy.lib.gf3_demap(vp(d), 0, vp(d), None, None) == 0
    assert gf3b200.launch_count() == n0
    assert phy.lib.gf3_rx_receive(phy._plan, vp(d), None, 1, None, vp(d), vp(d), vp(d), None, vp(d), 16, None, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_rx_receive(phy._plan, vp(d), None, 1, vp(phy.known), vp(d), vp(d), vp(d), None, vp(d), 6, None, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_eq_apply(phy._plan, None, 1, vp(d), vp(d), vp(d), vp(d), None, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_demap(None, 4, vp(d), None, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_tx_frame(phy._plan, vp(d), 1, None, 8, vp(phy.known), vp(d), None) == _lib.GF3_ERR_INVALID
    assert gf3b200.launch_count() == n0
    # bad arguments: error codes + message, no exception across the ABI, no launch
    assert phy.lib.gf3_rx_demod(phy._plan, vp(d), None, 1, vp(d), vp(d), vp(d), None, vp(d), 6, None, None) == _lib.GF3_ERR_INVALID
    assert b"bits_stride" in phy.lib.gf3_last_error()
    assert phy.lib.gf3_rx_demod(phy._plan, vp(d), None, 1, vp(d), vp(d), vp(d), None, None, 0, None, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_rx_estimate(phy._plan, None, None, 1, vp(phy.known), vp(d), vp(d), vp(d), None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_channel_sim(vp(d), 8, 1, 8, vp(d), 65, None, 0, vp(d), 8, None) == _lib.GF3_ERR_INVALID
    assert phy.lib.gf3_xcorr(phy._plan, vp(d), 8, 1, 8, vp(d), 10, vp(d), vp(d), None) == _lib.GF3_ERR_INVALID
    assert gf3b200.launch_count() == n0
    with pytest.raises(gf3b200.Gf3Error):
        gf3b200.Phy(N=1000, cp=16, lo=3, hi=100)


def test_tx_sync_outputs_guarded(known_sequence):
    """tx_modulate, xcorr and peak_pick write only inside their buffers (ragged sizes)."""
    torch = _torch()
    import gf3b200
    from gf3b200._lib import check
    phy = gf3b200.Phy(N=512, cp=10, lo=2, hi=255, n_pilots=1, packet_len=5, known_sequence=known_sequence, fit_lo=20, fit_hi=200)
    B, pk = 3, 2
    gen = torch.Generator(device="cuda").manual_seed(3)
    bits = torch.randint(0, 256, (B, pk, phy.bits_stride), dtype=torch.uint8, device="cuda", generator=gen)
    f = torch.randint(0, 4, (B, phy.K - phy.Nd), device="cuda", generator=gen)
    filler = (((1 - 2 * (f & 1)) + 1j * (1 - 2 * (f >> 1))) / np.sqrt(2)).to(torch.complex64).contiguous()
    T = phy.tx_len(pk)
    stride = T + 5
    g = Guarded(torch, B * stride * 4, torch.float32, (B, stride))
    g.view.fill_(7.0)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    check(phy.lib.gf3_tx_modulate(phy._plan, vp(bits), phy.bits_stride, vp(filler), vp(phy.known), B, pk, vp(g.view), stride, None))
    torch.cuda.synchronize()
    g.check("tx")
    assert bool(torch.all(g.view[:, T:] == 7.0)) and bool(torch.all(g.view[:, :T].abs() < 1.0))
    r = g.view[:, :T].contiguous()
    plen = T + phy.chirp_len - 1
    gP = Guarded(torch, B * plen * 4, torch.float32, (B, plen))
    gm = Guarded(torch, B * 4, torch.float32, (B,))
    work = torch.empty(int(phy.lib.gf3_xcorr_work_bytes(phy._plan, B, T)), dtype=torch.uint8, device="cuda")
    check(phy.lib.gf3_xcorr(phy._plan, vp(r), T, B, T, vp(gP.view), plen, vp(gm.view), vp(work), None))
    gk = Guarded(torch, B * 2 * 8, torch.int64, (B, 2))          # fewer slots than detections (3 per stream)
    gc = Guarded(torch, B * 4, torch.int32, (B,))
    gw = Guarded(torch, int(phy.lib.gf3_peak_pick_work_bytes(phy._plan, B, T)), torch.uint8, (-1,))
    check(phy.lib.gf3_peak_pick(phy._plan, vp(gP.view), plen, B, T, vp(gm.view), vp(gk.view), 2, vp(gc.view), vp(gw.view), None))
    torch.cuda.synchronize()
    for gg, name in ((gP, "P"), (gm, "pmax"), (gk, "peaks"), (gc, "count"), (gw, "peak_pick work")):
        gg.check(name)
    # the final chirp ends exactly at the end of r: the reference's wipe-out quirk gives 0 detections
    assert gc.view.cpu().tolist() == [0, 0, 0]
    r2 = torch.zeros((B, T + 4), device="cuda")
    r2[:, :T] = r
    P2, m2 = phy.xcorr(r2)
    peaks, cnt = phy.peak_pick(P2, m2, T + 4, 2)
    assert cnt.cpu().tolist() == [pk + 1] * B                    # count reports all, only max_peaks stored
    assert peaks[:, 0].cpu().tolist() == [phy.chirp_len - 2] * B


def test_receive_chain_replays_from_a_cuda_graph(known_sequence):
    """The whole receive chain is one asynchronous launch on the caller's stream with no host
    synchronisation inside the library: it can be captured once and replayed from a CUDA graph."""
    torch = _torch()
    import gf3b200
    from gf3b200 import synth
    phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=4, packet_len=24, known_sequence=known_sequence,
                      fit_lo=125, fit_hi=250)
    b = synth.make_batch(phy, 96, 1, snr_db=15.0, seed=21)
    sym = synth.packets_from_streams(phy, b).contiguous()
    n = sym.shape[0]
    flat = sym.reshape(-1)
    eager, Hs0, He0, sl0 = phy.rx_receive(flat, n, xor=True)
    torch.cuda.synchronize()
    out = torch.zeros_like(eager)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        phy.rx_receive(flat, n, xor=True, out=out)                      # warm-up on the capture stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            _, Hs, He, sl = phy.rx_receive(flat, n, xor=True, out=out)
    torch.cuda.current_stream().wait_stream(side)
    out.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager) and torch.equal(Hs, Hs0) and torch.equal(sl, sl0)
    # new samples in the same buffer: the replay sees them
    flat.mul_(-1.0)
    neg, _, _, _ = phy.rx_receive(flat, n, xor=True)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, neg)
