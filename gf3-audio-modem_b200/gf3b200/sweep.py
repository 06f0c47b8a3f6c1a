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
    """tx_modulate -> channel_sim -> xcorr / peak_pick / packet offsets -> rx_receive (estimate + data symbols)
    -> ber_count, all on the device and batched over an explicit list of stream ids.  A stream's bits,
    filler, channel and noise depend on (seed, stream id, SNR) only (synth.make_batch), so the counters
    do not depend on how the streams are sharded over ranks or chunks."""
    import torch
    from . import synth

    def count(stream_ids, snr_db):
        n = len(stream_ids)
        dev = phy.device
        b = synth.make_batch(phy, n, pk_per_stream, snr_db=snr_db, seed=seed, lead=64, trail=8, stream_ids=stream_ids)
        r, bits = b["r"], b["bits"]
        nbytes = (phy.bits_per_packet + 7) // 8
        fails = 0
        if use_sync:
            out = phy.receive_streams(r, pk_per_stream, xor=False)
            ok = out["ok"].bool()
            rx_bits = out["bits"].reshape(n, pk_per_stream, -1)
            fails = int((~ok).sum())
            if fails:      # streams whose chirps were not found: decode at the nominal position (counts as received garbage or luck)
                bad = torch.nonzero(~ok).reshape(-1)
                off = (b["starts"][bad] + (bad * r.stride(0))[:, None]).reshape(-1).contiguous()
                rx_bits[bad] = phy.rx_receive(r, len(bad) * pk_per_stream, off, xor=False)[0].reshape(len(bad), pk_per_stream, -1)
        else:
            off = (b["starts"] + (torch.arange(n, device=dev) * r.stride(0))[:, None]).reshape(-1).contiguous()
            rx_bits = phy.rx_receive(r, n * pk_per_stream, off, xor=False)[0].reshape(n, pk_per_stream, -1)
        cntr = torch.zeros(2, dtype=torch.int64, device=dev)
        a = rx_bits[:, :, :nbytes].contiguous()
        t = bits[:, :, :nbytes].contiguous()
        per_row = phy.bits_per_packet
        if per_row % 8 == 0:
            phy.ber_count(a, t, a.numel() * 8, cntr)
            e, nb = (int(v) for v in cntr.cpu())
        else:   # rows end inside a byte: pad bits are zero on both sides, so count bytes and report the true bit total
            phy.ber_count(a, t, a.numel() * 8, cntr)
            e = int(cntr[0])
            nb = n * pk_per_stream * per_row
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
    ap.add_argument("--chunk", type=int, default=512, help="streams per device batch")
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    phy = Phy(N=args.N, cp=args.cp, lo=args.lo, hi=args.hi, fit_lo=args.fit[0], fit_hi=args.fit[1])
    count_fn = make_gpu_count_fn(phy)
    count_fn(shard_streams(args.streams, rank, world)[:8], args.snr[0])          # warm-up: plan tables, allocator, first launches
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = sweep(count_fn, args.streams, args.snr, rank, world, dist if world > 1 else None,
                tensor_factory=lambda a: torch.tensor(a, dtype=torch.int64, device=phy.device), chunk=args.chunk)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=phy.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # device-timed, max over ranks
    if rank == 0:
        import hashlib
        print(json.dumps({"snr_db": args.snr, "bit_errors": res[:, 0].tolist(), "bits": res[:, 1].tolist(),
                          "sync_failures": res[:, 2].tolist(), "ber": (res[:, 0] / np.maximum(res[:, 1], 1)).tolist(),
                          "world_size": world, "streams": args.streams, "chunk": args.chunk, "seconds": float(ms) * 1e-3,
                          "stream_snr_points_per_s": args.streams * len(args.snr) / (float(ms) * 1e-3),
                          "counters_sha256": hashlib.sha256(np.ascontiguousarray(res).tobytes()).hexdigest(),
                          "geometry": dict(N=args.N, cp=args.cp, lo=args.lo, hi=args.hi, fit=args.fit)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
