#!/usr/bin/env python
"""Time the receive chain as two launches (gf3_rx_estimate + gf3_rx_demod) and as one (gf3_rx_receive)
on a bench workload.  usage: python tools/bench_chain.py [c3|c4|a2] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gf3-audio-modem_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
import gf3b200
from gf3b200 import synth


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    cfg, streams, desc = bench.WORKLOADS[wl]
    phy = gf3b200.Phy(**cfg)
    sym = torch.empty((streams, phy.pkt_samples), dtype=torch.float32, device=phy.device)
    for s0 in range(0, streams, 512):
        n = min(512, streams - s0)
        b = synth.make_batch(phy, n, 1, snr_db=20.0, seed=1234, first_stream=s0)
        sym[s0:s0 + n] = synth.packets_from_streams(phy, b)
        del b
    flat = sym.reshape(-1)
    out = torch.empty((streams, phy.bits_stride), dtype=torch.uint8, device=phy.device)

    def two():
        Hs, He, slope = phy.rx_estimate(flat, streams)
        phy.rx_demod(flat, streams, Hs, He, slope, xor=True, out=out)

    def one():
        phy.rx_receive(flat, streams, xor=True, out=out)

    res = {}
    for name, fn in (("two launches", two), ("one launch", one), ("two launches", two), ("one launch", one)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[name] = min(ms, res.get(name, 1e9))
        if name == "two launches":
            ref = out.clone()
        else:
            print("bits identical to the two-launch result:", bool(torch.equal(out, ref)))
    byt = bench.alg_bytes(cfg)[2] * streams
    peak = bench.peaks()[0]
    for k, v in res.items():
        print("%-13s %.4f ms/step  %.1f GB/s  %.1f %% of %.1f" % (k, v, byt / v / 1e6, 100 * byt / v / 1e6 / peak, peak))


if __name__ == "__main__":
    main()
