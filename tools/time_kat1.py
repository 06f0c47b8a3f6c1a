#!/usr/bin/env python
"""Wall time of the drop-in receiver on the reference's real recording (configs[0]: gr5ch1_signal.wav,
mode A2, XOR): receiver.receive(r) end to end from host numpy samples to host bits, per call.
The reference takes 14.8 s for the same call.  usage: python tools/time_kat1.py [reps]"""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "gf3-audio-modem_b200"))
import numpy as np
import torch

import OFDM


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    g = np.load(os.path.join(ROOT, "tests", "golden", "kat1_gr5ch1.npz"))
    r8 = g["wav_u8"]
    rx = OFDM.receiver(mode="A2", encoding="XOR")
    for name, sig in (("uint8 PCM as recorded", r8), ("float64 (r/1.0, as the notebook does)", r8 / 1.0)):
        ts = []
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                bits, Hs, He = rx.receive(sig)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ok = np.array_equal(np.packbits(bits.astype(np.uint8)), g["bits_packed"]) if "bits_packed" in g else None
        print("%-40s %d samples -> %d bits: first call %.1f ms, then median %.1f ms (min %.1f); equals the reference's bits: %s"
              % (name, len(sig), len(bits), ts[0] * 1e3, float(np.median(ts[1:])) * 1e3, min(ts[1:]) * 1e3, ok))


if __name__ == "__main__":
    main()
