#!/usr/bin/env python
"""Per-stage device timings of every C-ABI entry point on the C3 shape (CUDA events, device-resident
inputs): achieved algorithmic GB/s of each kernel group against the measured HBM peak.
usage: python tools/bench_stages.py [--streams 1024] [--N 1024 --cp 32 --lo 1 --hi 512]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf3-audio-modem_b200"))
import numpy as np
import torch
import gf3b200
from gf3b200 import synth


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--cp", type=int, default=32)
    ap.add_argument("--lo", type=int, default=1)
    ap.add_argument("--hi", type=int, default=512)
    ap.add_argument("--fit", type=int, nargs=2, default=[125, 250])
    a = ap.parse_args()
    peak = 6548.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    phy = gf3b200.Phy(N=a.N, cp=a.cp, lo=a.lo, hi=a.hi, fit_lo=a.fit[0], fit_hi=a.fit[1])
    B = a.streams
    batch = synth.make_batch(phy, B, 1, snr_db=20.0, seed=7, lead=96, trail=8)
    r, bits = batch["r"].contiguous(), batch["bits"]
    T = r.shape[1]
    K, Nd, P, L, sl, Lc = phy.K, phy.Nd, phy.P, phy.L, phy.symlen, phy.chirp_len
    gen = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randint(0, 4, (B, max(K - Nd, 1)), device="cuda", generator=gen)
    filler = (((1 - 2 * (f & 1)) + 1j * (1 - 2 * (f >> 1))) / np.sqrt(2)).to(torch.complex64).contiguous() if K > Nd else None
    out = {}

    txbuf = torch.empty((B, (phy.tx_len(1) + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: phy.tx_modulate(bits, filler, B, 1, out=txbuf))
    out["tx_modulate"] = (ms, B * (L * Nd // 4 + phy.tx_len(1) * 4))
    x = txbuf[:, : phy.tx_len(1)]
    taps = batch["taps"]
    sig = batch["sigma"]
    ms = timeit(lambda: phy.channel_sim(x, taps, sig, 3))
    out["channel_sim (30 taps + AWGN)"] = (ms, B * phy.tx_len(1) * 8)
    ms = timeit(lambda: phy.xcorr(r), iters=5)
    out["xcorr (overlap-save matched filter)"] = (ms, B * T * 8)           # ideal: read r, write P
    Pm, pmax = phy.xcorr(r)
    ms = timeit(lambda: phy.peak_pick(Pm, pmax, T, 8))
    out["peak_pick"] = (ms, B * Pm.shape[1] * 4)
    peaks, cnt = phy.peak_pick(Pm, pmax, T, 8)
    ok = int((cnt == 2).sum())
    off = ((peaks[:, :1] + 2) + (torch.arange(B, device="cuda") * T)[:, None]).reshape(-1).contiguous()
    flat = r.reshape(-1)
    ms = timeit(lambda: phy.rx_estimate(flat, B, off))
    out["rx_estimate (unaligned sync offsets)"] = (ms, B * (2 * P * sl * 4 + 2 * K * 8 + 8))
    Hs, He, slope = phy.rx_estimate(flat, B, off)
    ob = torch.empty((B, phy.bits_stride), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: phy.rx_demod(flat, B, Hs, He, slope, off, out=ob))
    out["rx_demod (unaligned sync offsets)"] = (ms, B * (L * sl * 4 + L * Nd // 4 + 2 * K * 8 + 8))
    nb = (phy.bits_per_packet + 7) // 8
    raw = phy.rx_demod(flat, B, Hs, He, slope, off, xor=False)
    cntr = torch.zeros(2, dtype=torch.int64, device="cuda")
    aa, bb = raw[:, :nb].contiguous(), bits[:, 0, :nb].contiguous()
    ms = timeit(lambda: phy.ber_count(aa, bb, aa.numel() * 8, cntr))
    out["ber_count"] = (ms, 2 * aa.numel())
    cntr.zero_()
    phy.ber_count(aa, bb, aa.numel() * 8, cntr)
    e, n = (int(v) for v in cntr.cpu())
    total_ms = sum(v[0] for k, v in out.items() if k.split()[0] in ("xcorr", "peak_pick", "rx_estimate", "rx_demod"))
    print("# stages on %d streams, N=%d CP=%d Nd=%d, T=%d samples/stream; sync found %d/%d; BER %.4g" % (B, a.N, a.cp, Nd, T, ok, B, e / max(n, 1)))
    print("| stage | ms | algorithmic GB | GB/s | of measured HBM peak (%.0f GB/s) |\n|---|---|---|---|---|" % peak)
    for k, (ms, by) in out.items():
        print("| %s | %.3f | %.3f | %.0f | %.1f %% |" % (k, ms, by / 1e9, by / ms / 1e6, 100 * by / ms / 1e6 / peak))
    sym = B * (2 * P + L)
    print("full receive chain from raw streams (xcorr + peak_pick + estimate + demod): %.3f ms -> %.1f M sym/s, %.1f Gbit/s"
          % (total_ms, sym / total_ms / 1e3, B * phy.bits_per_packet / total_ms / 1e6))


if __name__ == "__main__":
    main()
