"""Synthetic workloads of BASELINE.json / SURVEY 8(d), generated ON THE DEVICE with the package's
own transmit and channel kernels (setup code, never inside a timed region)."""
import numpy as np
import torch


def random_channels(n_streams, n_taps=30, decay=5.0, first_stream=0, stream_ids=None):
    """Per-stream multipath: h_k ~ N(0, exp(-k/decay)), unit energy, seed 1000 + stream id, with a
    dominant first tap so the matched-filter peak sits on tap 0."""
    ids = np.arange(first_stream, first_stream + n_streams) if stream_ids is None else np.asarray(stream_ids)
    h = np.empty((len(ids), n_taps), dtype=np.float32)
    k = np.arange(n_taps)
    for s, sid in enumerate(ids):
        rng = np.random.default_rng(1000 + int(sid))
        t = rng.normal(0.0, np.sqrt(np.exp(-k / decay)))
        t[0] = abs(t[0]) + 1.0
        h[s] = t / np.sqrt(np.sum(t * t))
    return h


def make_batch(phy, n_streams, pk_per_stream=1, snr_db=20.0, seed=1234, first_stream=0, lead=0, trail=2,
               n_taps=30, stream_ids=None):
    """bits -> tx_modulate -> per-stream FIR + AWGN.  Returns dict(r [B, T] float32 received
    streams, bits uint8 [B, pk, bits_stride] transmitted (encoded) packed bits, starts int64 [B, pk]
    packet start offsets inside each row of r, taps, sigma).  Every stream's bits, filler, channel and
    noise depend on (seed, stream id) only -- ids = first_stream .. first_stream + n_streams - 1 or the
    explicit stream_ids -- so any sharding of the streams over calls or ranks sees the same streams."""
    dev = phy.device
    ids_np = np.arange(first_stream, first_stream + n_streams) if stream_ids is None else np.asarray(stream_ids, dtype=np.int64)
    assert len(ids_np) == n_streams
    ids = torch.from_numpy(ids_np.astype(np.int64)).to(dev)
    nbytes = (phy.bits_per_packet + 7) // 8
    bits = torch.zeros((n_streams, pk_per_stream, phy.bits_stride), dtype=torch.uint8, device=dev)
    for j in range(pk_per_stream):
        phy.random_bytes(n_streams, nbytes, seed * 1000003 + j, ids, out=bits[:, j, :nbytes])
    if phy.bits_per_packet % 8:
        bits[:, :, nbytes - 1] &= (0xFF00 >> (phy.bits_per_packet % 8)) & 0xFF      # pad bits of the last byte are 0
    filler = None
    if phy.K > phy.Nd:
        f = phy.random_bytes(n_streams, phy.K - phy.Nd, seed * 1000003 + 999983, ids).to(torch.int64)
        filler = (((1 - 2 * (f & 1)) + 1j * (1 - 2 * ((f >> 1) & 1))) / np.sqrt(2)).to(torch.complex64).contiguous()
    tx = phy.tx_modulate(bits, filler, n_streams, pk_per_stream)
    T = tx.shape[1] + lead + trail
    x = torch.zeros((n_streams, (T + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :T]
    x[:, lead:lead + tx.shape[1]] = tx
    del tx
    taps = torch.from_numpy(random_channels(n_streams, n_taps, stream_ids=ids_np)).to(dev)
    c0 = lead + phy.chirp_len
    sigma = None
    if snr_db is not None:
        y0 = phy.channel_sim(x, taps, None, 0)
        power = y0[:, c0:c0 + phy.pkt_samples].pow(2).mean(dim=1)
        del y0
        sigma = torch.sqrt(power) * (10.0 ** (-float(snr_db) / 20.0))
    r = phy.channel_sim(x, taps, sigma, seed * 7919 + int(round(float(snr_db if snr_db is not None else 0.0) * 16)), stream_ids=ids)
    starts = c0 + torch.arange(pk_per_stream, device=dev, dtype=torch.int64)[None, :] * (phy.chirp_len + phy.pkt_samples)
    starts = starts.expand(n_streams, pk_per_stream).contiguous()
    return dict(r=r, bits=bits, starts=starts, taps=taps, sigma=sigma, ids=ids)


def packets_from_streams(phy, batch):
    """Slice the (already synchronised) packets out of the received streams: contiguous
    sym float32 [B*pk, (2P+L)(N+cp)] -- the `sym` array of SURVEY 8(d).  (make_batch puts every
    stream's packets at the same offsets.)"""
    r, starts = batch["r"], batch["starts"]
    B, pk = starts.shape
    out = torch.empty((B, pk, phy.pkt_samples), dtype=torch.float32, device=r.device)
    for j in range(pk):
        s0 = int(starts[0, j])
        out[:, j] = r[:, s0:s0 + phy.pkt_samples]
    return out.reshape(B * pk, phy.pkt_samples)
