#!/usr/bin/env python
"""Small invocations of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Sizes are tiny (the tools slow kernels down by orders of magnitude) but cover: fused receive with packets split
between persistent CTAs, the two-launch chain, raw-stream sync (both peak-picker paths), transmit, N = 1024 and 4096."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf3-audio-modem_b200"))
import torch
import gf3b200
from gf3b200 import synth


def run(cfg, streams):
    phy = gf3b200.Phy(**cfg)
    b = synth.make_batch(phy, streams, 1, snr_db=15.0, seed=3, lead=37, trail=9)
    out = phy.receive_streams(b["r"], 1, xor=True)
    sym = synth.packets_from_streams(phy, b).contiguous()
    one = phy.rx_receive(sym.reshape(-1), streams, xor=True)[0]
    Hs, He, sl = phy.rx_estimate(sym.reshape(-1), streams)
    two = phy.rx_demod(sym.reshape(-1), streams, Hs, He, sl, xor=True)
    torch.cuda.synchronize()
    ok = int(out["ok"].sum())
    print("N=%d: %d streams, %d synchronised, fused == two-launch rows: %d / %d" % (cfg["N"], streams, ok, int((one == two).all(dim=1).sum()), streams))


if __name__ == "__main__":
    torch.cuda.set_device(0)
    run(dict(N=1024, cp=32, lo=1, hi=512, n_pilots=4, packet_len=40, fit_lo=125, fit_hi=250), 12)
    run(dict(N=4096, cp=224, lo=100, hi=1500, n_pilots=2, packet_len=9), 3)
    run(dict(N=256, cp=16, lo=3, hi=100, n_pilots=2, packet_len=33, fit_lo=10, fit_hi=90), 700)      # > 2 waves of streams: one-kernel peak picker
    print("sanitize smoke done, launches:", gf3b200.launch_count())
