"""Stream-sharded tx -> channel -> rx BER sweep over SNR (BASELINE.json configs[4]).

Every stream is an independent unit (its own bits, channel, noise, sync and channel estimate), so
rank r of W owns the streams with id % W == r and the data path has NO collective.  The only
exchange is one all-reduce(SUM) of the int64 counters [n_snr, 3] = (bit errors, bits, streams whose
chirp sync failed) at the end -- NCCL over NVLink on the GPUs, gloo in the CPU tests.

    torchrun --nproc-per-node 8 -m gf3b200.sweep --streams 4096 --snr 0 2 4 ... 20
"""
import argparse
import json
import os
import sys

import numpy as np

if __package__ in (None, ""):          # run as a script (torchrun path/to/sweep.py): make relative imports work
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gf3b200  # noqa: F401
    __package__ = "gf3b200"


def shard_streams(n_streams, rank, world):
    """Stream ids owned by `rank`: id % world == rank."""
    return np.arange(rank, n_streams, world, dtype=np.int64)


def reduce_counters(local, dist=None):
    """all-reduce(SUM) of the local int64 counters; identity when not distributed."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    return local


def sweep(count_fn, n_streams, snrs_db, rank=0, world=1, dist=None, tensor_factory=None, chunk=512):
    """Generic sharded sweep.  count_fn(stream_ids, snr_db) -> (errors, bits, sync_failures) for the
    given streams; returns the globally reduced counters as an int64 array [n_snr, 3]."""
    import torch
    mine = shard_streams(n_streams, rank, world)
    make = tensor_factory or (lambda a: torch.tensor(a, dtype=torch.int64))
    rows = []
    for snr in snrs_db:
        e = b = f = 0
        for c0 in range(0, len(mine), chunk):
            de, db, df = count_fn(mine[c0:c0 + chunk], snr)
            e, b, f = e + int(de), b + int(db), f + int(df)
        rows.append([e, b, f])
    t = make(rows)
    reduce_counters(t, dist)
    return t.cpu().numpy()


def make_gpu_count_fn(phy, pk_per_stream=1, seed=1234, use_sync=True):
    """tx_modulate -> channel_sim -> xcorr/peak_pick -> rx_receive (estimate + data symbols) -> ber_count on the
    device, for an explicit list of stream ids."""
    import torch
    from . import synth

    def count(stream_ids, snr_db):
        n = len(stream_ids)
        dev = phy.device
        # per-stream reproducible bits: seed derived from the stream id, independent of sharding
        bits = torch.empty((n, pk_per_stream, phy.bits_stride), dtype=torch.uint8, device=dev)
        fill = torch.empty((n, max(phy.K - phy.Nd, 1)), dtype=torch.int64, device=dev)
        for i, sid in enumerate(stream_ids):
            g = torch.Generator(device=dev).manual_seed(seed * 1000003 + int(sid))
            bits[i] = torch.randint(0, 256, (pk_per_stream, phy.bits_stride), dtype=torch.uint8, device=dev, generator=g)
            fill[i] = torch.randint(0, 4, (fill.shape[1],), device=dev, generator=g)
        nbytes = (phy.bits_per_packet + 7) // 8
        bits[:, :, nbytes:] = 0
        filler = None
        if phy.K > phy.Nd:
            filler = (((1 - 2 * (fill & 1)) + 1j * (1 - 2 * (fill >> 1))) / np.sqrt(2)).to(torch.complex64).contiguous()
        tx = phy.tx_modulate(bits, filler, n, pk_per_stream)
        lead, trail = 64, 8
        T = tx.shape[1] + lead + trail
        x = torch.zeros((n, (T + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :T]
        x[:, lead:lead + tx.shape[1]] = tx
        taps = torch.from_numpy(synth.random_channels(n, stream_ids=stream_ids)).to(dev)
        y0 = phy.channel_sim(x, taps, None, 0)
        c0 = lead + phy.chirp_len
        sigma = torch.sqrt(y0[:, c0:c0 + phy.pkt_samples].pow(2).mean(dim=1)) * (10.0 ** (-float(snr_db) / 20.0))
        seeds = int(seed) * 7919 + int(round(float(snr_db) * 16))
        # noise must not depend on which rank / chunk a stream lands in: one launch per stream id
        r = torch.empty_like(y0)
        for i, sid in enumerate(stream_ids):
            r[i:i + 1] = phy.channel_sim(x[i:i + 1], taps[i:i + 1], sigma[i:i + 1], seeds * 65537 + int(sid))
        r = r.contiguous()
        known_starts = c0 + torch.arange(pk_per_stream, device=dev, dtype=torch.int64) * (phy.chirp_len + phy.pkt_samples)
        starts = known_starts[None, :].expand(n, pk_per_stream).clone()
        fails = 0
        if use_sync:
            P, pmax = phy.xcorr(r)
            peaks, cnt = phy.peak_pick(P, pmax, r.shape[1], pk_per_stream + 4)
            ok = cnt == pk_per_stream + 1
            det = peaks[:, :pk_per_stream] + 2
            starts = torch.where(ok[:, None], det, starts)
            starts = torch.minimum(starts, torch.tensor(r.shape[1] - phy.pkt_samples, device=dev))
            fails = int((~ok).sum())
        off = (starts + (torch.arange(n, device=dev) * r.shape[1])[:, None]).reshape(-1).contiguous()
        out = phy.rx_receive(r.reshape(-1), n * pk_per_stream, off, xor=False)[0]     # estimate + data symbols, one launch
        cntr = torch.zeros(2, dtype=torch.int64, device=dev)
        a = out[:, :nbytes].contiguous()
        b = bits.reshape(n * pk_per_stream, -1)[:, :nbytes].contiguous()
        phy.ber_count(a, b, a.numel() * 8, cntr)
        e, nb = (int(v) for v in cntr.cpu())
        return e, nb, fails

    return count


def main(argv=None):
    import torch
    import torch.distributed as dist
    from . import Phy
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--snr", type=float, nargs="+", default=[float(s) for s in range(0, 22, 2)])
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--cp", type=int, default=32)
    ap.add_argument("--lo", type=int, default=1)
    ap.add_argument("--hi", type=int, default=512)
    ap.add_argument("--fit", type=int, nargs=2, default=[125, 250])
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    phy = Phy(N=args.N, cp=args.cp, lo=args.lo, hi=args.hi, fit_lo=args.fit[0], fit_hi=args.fit[1])
    res = sweep(make_gpu_count_fn(phy), args.streams, args.snr, rank, world, dist if world > 1 else None,
                tensor_factory=lambda a: torch.tensor(a, dtype=torch.int64, device=phy.device), chunk=64)
    if rank == 0:
        print(json.dumps({"snr_db": args.snr, "bit_errors": res[:, 0].tolist(), "bits": res[:, 1].tolist(),
                          "sync_failures": res[:, 2].tolist(), "ber": (res[:, 0] / np.maximum(res[:, 1], 1)).tolist(),
                          "world_size": world, "streams": args.streams}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
