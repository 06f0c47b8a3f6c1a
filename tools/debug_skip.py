import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf3-audio-modem_b200"))
import torch, gf3b200
from gf3b200 import synth
phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, fit_lo=125, fit_hi=250)
for snr in (20.0, 8.0):
    b = synth.make_batch(phy, 1024, 1, snr_db=snr, seed=1234, lead=700, trail=1308)
    r = b["r"]
    B, T = r.shape
    nblk = (T + phy.chirp_len - 1 + 2047) // 2048
    phy.sync_streams(r, 4, detect_only=True)
    torch.cuda.synchronize()
    bm = phy._last_sync_work[: B * nblk * 4].view(torch.float32).reshape(B, nblk)
    skipped = torch.isinf(bm) & (bm < 0)
    per_stream = skipped.float().mean(dim=1)
    print("snr", snr, "blocks/stream", nblk, "skipped overall %.3f" % float(skipped.float().mean()), "per-stream quantiles", [round(float(per_stream.quantile(q)), 3) for q in (0.0, 0.1, 0.25, 0.5, 0.75, 1.0)],
          "streams with <50%% skipped: %d" % int((per_stream < 0.5).sum()))
    P, pmax, _, _ = phy.sync_streams(r, 4)
    # bound quality: (true) block maxima relative to pmax
    bm2 = phy._last_sync_work[: B * nblk * 4].view(torch.float32).reshape(B, nblk)
    print("   dense: median blockmax/pmax %.4f" % float((bm2 / pmax[:, None]).median()))
