#!/usr/bin/env python
"""Per-stage times of one receive of the real recording (gr5ch1_signal.wav, mode A2) through the Phy
calls the drop-in makes; run under `ncu --metrics gpu__time_duration.sum` for the kernel times.
usage (from the repo root): python tools/kat1_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gf3-audio-modem_b200"))
import numpy as np, torch, gf3b200
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "kat1_gr5ch1.npz"))
r8 = g["wav_u8"]
phy = gf3b200.Phy(N=4096, cp=224, lo=100, hi=1500)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    t0 = time.perf_counter()
    d8 = torch.from_numpy(np.ascontiguousarray(r8).reshape(1, -1)).cuda()
    e = [ev() for _ in range(6)]
    e[0].record(); d_r = phy.pcm_to_f32(d8)
    e[1].record(); P, pmax = phy.xcorr(d_r)
    e[2].record(); peaks, count = phy.peak_pick(P, pmax, d_r.shape[1], 8)
    e[3].record(); n = int(count[0].item()); pk = peaks[0, :n].cpu().numpy()
    starts = torch.from_numpy((pk + 2)[:-1].astype(np.int64)).cuda()
    e[4].record(); out = phy.rx_receive(d_r.reshape(-1), len(pk) - 1, starts, xor=True)
    e[5].record(); bits = phy.unpack_bits(out[0]); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("wall %.2f ms | pcm %.3f xcorr %.3f peak_pick %.3f host-roundtrip %.3f rx_receive %.3f ms" % ((t1 - t0) * 1e3, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), e[4].elapsed_time(e[5])))
