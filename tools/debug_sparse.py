import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf3-audio-modem_b200"))
import torch, gf3b200
from gf3b200 import synth
phy = gf3b200.Phy(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, fit_lo=125, fit_hi=250)
B = 160
for snr in (3.0, 12.0, 6.0):
    b = synth.make_batch(phy, B, 1, snr_db=snr, seed=int(100 + snr), lead=513, trail=3)
    r = b["r"]
    T = r.shape[1]
    nblk = (T + phy.chirp_len - 1 + 2047) // 2048
    P, pmax, peaks, count = phy.sync_streams(r, 16)
    torch.cuda.synchronize()
    bmd = phy._last_sync_work[: B * nblk * 4].view(torch.float32).reshape(B, nblk).clone()
    for rep in range(6):
        P2, pmax2, peaks2, count2 = phy.sync_streams(r, 16, detect_only=True)
        torch.cuda.synchronize()
        bms = phy._last_sync_work[: B * nblk * 4].view(torch.float32).reshape(B, nblk).clone()
        bad = torch.nonzero((count != count2) | (peaks != peaks2).any(dim=1) | (pmax != pmax2)).reshape(-1).tolist()
        if not bad:
            continue
        for s in bad[:2]:
            comp = ~(torch.isinf(bms[s]) & (bms[s] < 0))
            hot = bmd[s] / pmax[s] > 0.4
            print("snr", snr, "rep", rep, "stream", s, "count", int(count[s]), int(count2[s]), "peaks", peaks[s, :3].tolist(), peaks2[s, :3].tolist(), "pmax", float(pmax[s]), float(pmax2[s]))
            print("   hot blocks (dense):", torch.nonzero(hot).reshape(-1).tolist(), "computed in sparse:", torch.nonzero(comp).reshape(-1).tolist()[:40])
            print("   blockmax mismatch on computed blocks:", torch.nonzero(comp & (bms[s] != bmd[s])).reshape(-1).tolist()[:10])
            for hb in torch.nonzero(hot).reshape(-1).tolist():
                lo, hi = max(0, hb * 2048 - 2), min(P.shape[1], (hb + 1) * 2048 + 2)
                d = (P[s, lo:hi] != P2[s, lo:hi])
                print("   hot block", hb, "P differs at", (torch.nonzero(d).reshape(-1) + lo).tolist()[:10], "computed prev/this/next", [bool(comp[x]) for x in (hb - 1, hb, min(hb + 1, nblk - 1))])
        break
    else:
        print("snr", snr, "all 6 repetitions identical")
