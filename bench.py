#!/usr/bin/env python
"""bench.py -- headline benchmark of the GF3 B200 physical layer (BASELINE.json metric:
demodulated Mbit/s and OFDM symbols/s, % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c3-raw|c4|c4-long|a2|a2-raw|w2048] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.
  * packet workloads (c3, c4, c4-long, a2, w2048): the receive chain (gf3_rx_receive: channel estimate +
    data symbols) over pre-sliced received packets.  Default = BASELINE.json configs[2] ("C3": 4096
    independent streams x 1 packet, N=1024, CP=32, 511 data bins, P=20, L=180, random 30-tap multipath +
    AWGN 20 dB) -- the configuration the metric is quoted on;
  * c3-raw: the same streams as raw audio (chirp | packet | chirp with a lead-in), receiver.receive as
    the reference runs it: matched filter -> detection rule -> packet offsets -> receive chain
    (gf3_xcorr, gf3_peak_pick, gf3_peaks_to_offsets, gf3_rx_receive).
configs[0]/[1] are single-stream decodes and are covered as parity tests (KAT-1, KAT-4).  One process
per GPU (torchrun for N > 1): streams are sharded across ranks with no data-path collective (weak
scaling); NCCL all-reduces only the bit-error counters.  Prints ONE JSON line on rank 0.

After the timed region a sample of the ACTUAL timed batch goes through the float64 oracle and the
line reports `parity` (bits of the timed output vs the oracle's, decisions within 1e-5 of a boundary
counted separately, constellation error) -- the device-resident result and the end-to-end result.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "gf3-audio-modem_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# fit window: the reference hard-codes carrier indices [500:1000) for K=2047 (OFDM.py:462); at K=511 that
# clips to 11 band-edge bins and the drift estimate is noise, so the N=1024 workloads use the same band
# fraction, [125:250) ("BER-quality sweeps may expose fit_lo/fit_hi"; NOT the reference's literal
# parameter -- the literal window at this shape is parity-tested in tests/test_gpu_scale_parity.py)
C3 = dict(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, fit_lo=125, fit_hi=250)
WORKLOADS = {
    # name: (params, streams per GPU, raw streams?, description)
    "c3": (C3, 4096, False,
           "C3 (BASELINE.json configs[2]): 4096 streams x 1 packet, N=1024, CP=32, Nd=511, P=20, L=180, fit window [125:250) (not the reference's literal [500:1000]), random 30-tap multipath + AWGN 20 dB"),
    "c3-raw": (C3, 2048, True,
               "C3-raw (BASELINE.json configs[2] as raw audio): 2048 streams of [lead-in | chirp | packet | chirp | tail], N=1024, CP=32, Nd=511, P=20, L=180, chirp 5280, fit window [125:250), random 30-tap multipath + AWGN 20 dB; step = matched filter + detection + receive chain"),
    "c4": (dict(N=4096, cp=704, lo=1, hi=2047, n_pilots=20, packet_len=180), 512, False,
           "C4 (BASELINE.json configs[3], mode B1): 512 streams x 1 packet, N=4096, CP=704, Nd=2046, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
    "c4-long": (dict(N=4096, cp=704, lo=1, hi=2047, n_pilots=20, packet_len=1440), 64, False,
                "C4 long frames (BASELINE.json configs[3], mode B1): 64 streams x 1 packet, N=4096, CP=704, Nd=2046, P=20, L=1440, random 30-tap multipath + AWGN 20 dB"),
    "w2048": (dict(N=2048, cp=64, lo=1, hi=1024, n_pilots=20, packet_len=180, fit_lo=250, fit_hi=500), 2048, False,
              "W2048 (parity-test geometry): 2048 streams x 1 packet, N=2048, CP=64, Nd=1023, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
    "a2": (dict(N=4096, cp=224, lo=100, hi=1500, n_pilots=20, packet_len=180), 512, False,
           "A2 (mode of the real recording): 512 streams x 1 packet, N=4096, CP=224, Nd=1400, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
    "a2-raw": (dict(N=4096, cp=224, lo=100, hi=1500, n_pilots=20, packet_len=180), 256, True,
               "A2-raw (mode of the real recording, as raw audio): 256 streams of [lead-in | chirp | packet | chirp | tail], N=4096, CP=224, Nd=1400, P=20, L=180, chirp 21600 (11 partitions), random 30-tap multipath + AWGN 20 dB; step = matched filter + detection + receive chain"),
}
METRIC, UNIT = "demodulated_mbit_per_s", "Mbit/s"
RAW_LEAD_MAX, RAW_TAIL = 2000, 8


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes(cfg):
    """Algorithmic bytes per packet: demod kernel, estimate kernel, whole chain."""
    N, cp, P, L = cfg["N"], cfg["cp"], cfg["n_pilots"], cfg["packet_len"]
    K, Nd = N // 2 - 1, cfg["hi"] - cfg["lo"]
    demod = L * (N + cp) * 4 + L * Nd * 2 // 8 + 2 * K * 8 + 8
    est = 2 * P * (N + cp) * 4 + 2 * K * 8 + 8
    chain = (2 * P + L) * (N + cp) * 4 + L * Nd * 2 // 8 + 2 * K * 8 + 4
    return demod, est, chain


def static_config(desc, streams, world, gb):
    """The same dict in both arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": desc, "streams_per_gpu": streams, "packets_per_step": world * streams,
            "l2_policy": "input batch (%.2f GB per GPU) is larger than the 126 MB L2; no flush needed" % gb}


def oracle_params(cfg):
    from oracle import gf3_oracle as orc
    import gf3b200.phy as _phy
    return orc.Params(N=cfg["N"], cp=cfg["cp"], lo=cfg["lo"], hi=cfg["hi"], n_pilots=cfg["n_pilots"],
                      packet_len=cfg["packet_len"], known_sequence=_phy.default_known_sequence(), encoding="XOR",
                      fit_lo=cfg.get("fit_lo", 500), fit_hi=cfg.get("fit_hi", 1000))


# ------------------------------------------------------------------------------ CPU arms
def _cpu_streams(cfg, n, seed=99, raw=False):
    """Synthetic received packets (or raw streams) of the workload, made on the CPU with the oracle's
    transmit chain + scipy FIR + AWGN: float64 [n, 2P+L, N+cp], or a list of n raw streams."""
    import numpy as np
    from scipy.signal import lfilter
    from oracle import gf3_oracle as orc
    p = oracle_params(cfg)
    rng = np.random.default_rng(seed)
    out = [] if raw else np.empty((n, p.syms_per_packet, p.sym_len))
    bits_all = []
    k = np.arange(30)
    for i in range(n):
        bits = rng.integers(0, 2, p.data_bits_per_symbol * p.packet_len)
        tx = orc.transmit(p, bits, rng=_RngShim(rng))
        t = rng.normal(0.0, np.sqrt(np.exp(-k / 5.0)))
        t[0] = abs(t[0]) + 1.0
        t /= np.sqrt(np.sum(t * t))
        lead = int(rng.integers(0, RAW_LEAD_MAX)) if raw else 0
        y = lfilter(t, 1.0, np.concatenate([np.zeros(lead), tx, np.zeros(RAW_TAIL if raw else 0)]))
        c0 = lead + p.chirp_length
        sg = np.sqrt(np.mean(y[c0:c0 + p.packet_samples] ** 2)) * 0.1                    # 20 dB
        y = y + rng.normal(0, sg, len(y))
        if raw:
            out.append(y)
        else:
            out[i] = y[c0:c0 + p.packet_samples].reshape(p.syms_per_packet, p.sym_len)
        bits_all.append(bits)
    return p, out, bits_all


class _RngShim:
    """Gives a numpy Generator the two legacy method names the oracle's transmit() draws with."""

    def __init__(self, g):
        self.g = g

    def binomial(self, n, p, size):
        return self.g.binomial(n, p, size)

    def choice(self, a, size, replace=True):
        return self.g.choice(a, size=size, replace=replace)


_CPU_JOB = None      # (params, input, raw) inherited by the forked workers: nothing is pickled per step


def _cpu_receive(p, rx, raw):
    from oracle import gf3_oracle as orc
    if raw:
        return sum(len(orc.receive(p, r)["bits"]) for r in rx)
    return len(orc.receive_symbols(p, rx)["bits"])


def _cpu_worker(_):
    p, rx, raw = _CPU_JOB
    return _cpu_receive(p, rx, raw)


def cpu_baseline_single_thread(cfg, raw, budget_s=12.0):
    """The numpy oracle (float64 port of the reference's receive path) on one host thread over a
    bounded sample of the workload."""
    n0 = 8 if raw else 16
    p, rx, bits = _cpu_streams(cfg, n0, raw=raw)
    _cpu_receive(p, rx[:2], raw)                              # warm-up
    t0 = time.perf_counter()
    nbits = npk = 0
    while time.perf_counter() - t0 < budget_s:
        nbits += _cpu_receive(p, rx, raw)
        npk += n0
    dt = time.perf_counter() - t0
    what = "raw streams (oracle.receive: scipy matched filter + detection rule + receive chain)" if raw else "packets"
    return dict(value=nbits / dt / 1e6, unit=UNIT, cores=1, kind="port",
                sample="%d %s (%d OFDM symbols) of the workload in %.1f s, float64 numpy oracle (oracle/gf3_oracle.py), 1 thread"
                       % (npk, what, npk * p.syms_per_packet, dt),
                symbols_per_s=npk * p.syms_per_packet / dt)


def run_reference_arm(args, cfg, streams, raw, desc):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself is
    Python under /root/reference and cannot travel to the GPU box, so this times the numpy oracle
    port (pinned to the reference: tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py) on all
    host cores, one process per core, each step a bounded sample of the workload."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    per_core = 4 if raw else 16
    global _CPU_JOB
    p, rx, _ = _cpu_streams(cfg, per_core, raw=raw)
    _CPU_JOB = (p, rx, raw)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        work = list(range(cores))
        for _ in range(args.warmup):
            pool.map(_cpu_worker, work)
        t0 = time.perf_counter()
        nbits = 0
        for _ in range(args.steps):
            nbits += sum(pool.map(_cpu_worker, work))
        dt = time.perf_counter() - t0
    value = nbits / dt / 1e6
    npk = cores * per_core * args.steps
    gb = streams * (p.packet_samples + (2 * p.chirp_length + RAW_LEAD_MAX + RAW_TAIL if raw else 0)) * 4 / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(desc, streams, world, gb),
        "symbols_per_s": npk * p.syms_per_packet / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d %s per step (%d per core x %d cores) x %d steps, one process per host core, float64 numpy oracle port of OFDM.py:%s"
                                   % (cores * per_core, "raw streams" if raw else "packets", per_core, cores, args.steps,
                                      "356-372,391-609" if raw else "391-609")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.enabled = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:     # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            if self.enabled:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0,
                    "note": "NVML unavailable" if not self.ok else "no samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------ parity of what was timed
def parity_check(phy, cfg, samples_dev, timed_bits, raw, n_check, extra_results=()):
    """Run the float64 oracle on a sample of the ACTUAL timed batch and compare the TIMED output with it.
    samples_dev: [n, pkt_samples] packets or [n, T] raw streams (device); timed_bits [n, bits_stride] (device
    or host).  extra_results: (name, bits [n, bits_stride], samples or None) of further paths (end to end)."""
    import numpy as np
    import torch
    from oracle import gf3_oracle as orc
    from oracle.parity import classify_bit_diffs
    p = oracle_params(cfg)
    n = samples_dev.shape[0]
    idx = np.unique(np.linspace(0, n - 1, min(n_check, n)).round().astype(np.int64))
    didx = torch.from_numpy(idx).to(samples_dev.device)
    dc = p.data_carriers - 1
    nbytes = (phy.bits_per_packet + 7) // 8

    def oracle_on(sub64):
        if raw:
            outs = []
            for r in sub64:
                try:
                    outs.append(orc.receive(p, r, want_eq=True))
                except ValueError:      # fewer than two detections: the reference's vstack / reshape raises (OFDM.py:400-403)
                    outs.append(dict(starts=[], peaks=np.flatnonzero(orc.chirp_method(p, r))))
            return outs
        return orc.receive_symbols(p, sub64.reshape(len(sub64), p.syms_per_packet, p.sym_len), want_eq=True)

    def compare(bits_rows, ref, eq_gpu=None, peaks_gpu=None):
        got = np.unpackbits(np.asarray(bits_rows)[:, :nbytes], axis=1)[:, : phy.bits_per_packet]
        if raw:
            sync_ok = [len(o["starts"]) == 1 for o in ref]
            keep = [i for i, o in enumerate(ref) if sync_ok[i]]
            ref_bits = np.concatenate([ref[i]["bits"] for i in keep]) if keep else np.zeros(0, np.int64)
            ref_eq = np.concatenate([ref[i]["eq"][:, dc] for i in keep]) if keep else np.zeros((0, len(dc)), complex)
            got = got[keep]
            res = classify_bit_diffs(got.reshape(-1), ref_bits, ref_eq) if keep else dict(n_bits=0, n_diff=0, near_1e5=0, near_scaled=0, within_eq_tol=0, beyond=0, worst_margin=0.0, n_points_near_1e5=0)
            res["streams_oracle_sync_ok"] = int(sum(sync_ok))
            if peaks_gpu is not None:
                res["sync_index_mismatches"] = int(sum(not np.array_equal(o["peaks"], pg) for o, pg in zip(ref, peaks_gpu)))
        else:
            ref_bits, ref_eq = ref["bits"], ref["eq"][:, dc]
            res = classify_bit_diffs(got.reshape(-1), ref_bits, ref_eq)
        if eq_gpu is not None and len(ref_eq):
            rel = np.abs(eq_gpu - ref_eq) / np.maximum(np.abs(ref_eq), 1e-30)
            res["max_rel_eq_err"] = float(rel.max())
            res["p9999_rel_eq_err"] = float(np.quantile(rel, 0.9999))
        return res

    sub = samples_dev[didx]
    sub64 = sub.cpu().numpy().astype(np.float64)
    t0 = time.perf_counter()
    ref = oracle_on(sub64)
    t_or = time.perf_counter() - t0
    # constellation of the same packets (untimed launch with the constellation output on)
    peaks_gpu = None
    if raw:
        out = phy.receive_streams(sub.contiguous(), 1, xor=True, want_eq=True)
        cnt = out["count"].cpu().numpy()
        pk = out["peaks"].cpu().numpy()
        peaks_gpu = [pk[i, : cnt[i]] for i in range(len(idx))]
        eq_all = out["eq"].cpu().numpy().reshape(len(idx), -1, phy.K)[:, :, dc]
        keep = [i for i, o in enumerate(ref) if len(o["starts"]) == 1]
        eq_gpu = eq_all[keep].reshape(-1, len(dc)) if keep else None
    else:
        (_, eq), _, _, _ = phy.rx_receive(sub.reshape(-1), len(idx), xor=True, want_eq=True)
        eq_gpu = eq.cpu().numpy().reshape(-1, phy.K)[:, dc]
    tb = timed_bits[didx] if isinstance(timed_bits, torch.Tensor) and timed_bits.is_cuda else timed_bits[torch.from_numpy(idx)]
    res = compare(tb.cpu().numpy(), ref, eq_gpu, peaks_gpu)
    res.update(packets=int(len(idx)), oracle_seconds=round(t_or, 2),
               what="timed device-resident output vs oracle.%s (float64) on the same samples" % ("receive" if raw else "receive_symbols"))
    res["bit_mismatches"] = res.pop("n_diff")
    for name, bits_rows, samples in extra_results:
        r2 = ref if samples is None else oracle_on(samples[didx.to(samples.device)].cpu().numpy().astype(np.float64))
        rr = compare(bits_rows[torch.from_numpy(idx)].cpu().numpy(), r2)
        rr["bit_mismatches"] = rr.pop("n_diff")
        res[name] = rr
    return res


# ------------------------------------------------------------------------------ GPU arm
def make_input(phy, streams, rank, raw):
    """Synthetic input generated on the device by the package's own tx + channel kernels (set-up, untimed)."""
    import numpy as np
    import torch
    from gf3b200 import synth
    gen_chunk = 256
    tx_bits = torch.empty((streams, phy.bits_stride), dtype=torch.uint8, device=phy.device)
    if raw:
        T = phy.tx_len(1) + RAW_LEAD_MAX + RAW_TAIL
        Ts = (T + 3) // 4 * 4
        data = torch.empty((streams, Ts), dtype=torch.float32, device=phy.device)[:, :T]
        rng = np.random.default_rng(4321 + rank)
    else:
        data = torch.empty((streams, phy.pkt_samples), dtype=torch.float32, device=phy.device)
    for s0 in range(0, streams, gen_chunk):
        n = min(gen_chunk, streams - s0)
        if raw:
            lead = int(rng.integers(0, RAW_LEAD_MAX))                      # one random lead-in per chunk of streams
            b = synth.make_batch(phy, n, 1, snr_db=20.0, seed=1234, first_stream=rank * streams + s0, lead=lead,
                                 trail=T - phy.tx_len(1) - lead)
            data[s0:s0 + n] = b["r"]
        else:
            b = synth.make_batch(phy, n, 1, snr_db=20.0, seed=1234, first_stream=rank * streams + s0)
            data[s0:s0 + n] = synth.packets_from_streams(phy, b)
        tx_bits[s0:s0 + n] = b["bits"][:, 0]
        del b
    torch.cuda.synchronize()
    return data, tx_bits


def run_gpu_arm(args, cfg, streams, raw, desc):
    import numpy as np
    import torch
    import torch.distributed as dist
    import gf3b200
    from gf3b200.host import HostReceiver, bind_to_gpu_numa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)             # before any pinned allocation: host buffers land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    phy = gf3b200.Phy(**cfg)
    data, tx_bits = make_input(phy, streams, rank, raw)
    n_packets = streams
    out_bits = torch.empty((n_packets, phy.bits_stride), dtype=torch.uint8, device=phy.device)
    flat = data if raw else data.reshape(-1)
    T = data.shape[1]
    fused = phy.fused_receive       # one launch for the whole chain (N <= 2048), else estimate + demod
    NEV = 5
    state = {}

    def step(ev=None):
        rec = (lambda i: ev[i].record()) if ev is not None else (lambda i: None)
        rec(0)
        if raw:
            if args.split_sync:          # the two entry points separately (stage timings of matched filter / detection walk)
                P, pmax = phy.xcorr(data)
                rec(1)
                pk, cnt = phy.peak_pick(P, pmax, T, 4)
            else:                        # gf3_sync_streams: matched filter + detection (block maxima let the walk skip most of P)
                rec(1)
                _, _, pk, cnt = phy.sync_streams(data, 4, detect_only=not args.dense_sync)
            off, ok = phy.peaks_to_offsets(pk, cnt, data.stride(0), T, 1)
            rec(2)
            (phy.rx_receive_pcm if phy.staged_streams else phy.rx_receive)(data, n_packets, off, xor=True, out=out_bits)
            rec(3)
            rec(4)
            state["ok"], state["peaks"], state["count"] = ok, pk, cnt
        else:
            rec(1)
            rec(2)
            if fused:
                phy.rx_receive(flat, n_packets, xor=True, out=out_bits)             # XOR decode fused (Final System Test uses encoding="XOR")
                rec(3)
            else:
                Hs, He, slope = phy.rx_estimate(flat, n_packets)
                rec(3)
                phy.rx_demod(flat, n_packets, Hs, He, slope, xor=True, out=out_bits)
            rec(4)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(NEV)] for _ in range(args.steps)]
    l0 = gf3b200.launch_count()
    sampler.enabled = True
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(ev[k])
    t_end.record()
    barrier()
    sampler.enabled = False
    launches = gf3b200.launch_count() - l0
    elapsed_ms = t_start.elapsed_time(t_end)
    seg = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / args.steps for i in range(NEV - 1)]
    t = torch.tensor([elapsed_ms] + seg, dtype=torch.float64, device=phy.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, seg = float(t[0]), [float(x) for x in t[1:].cpu()]

    # ---- correctness of what was timed: BER against the transmitted bits (NCCL sum of counters)
    cnt = torch.zeros(3, dtype=torch.int64, device=phy.device)
    nbytes = (phy.bits_per_packet + 7) // 8
    if raw:
        raw_bits = phy.receive_streams(data, 1, xor=False)["bits"]
        cnt[2] = int((state["ok"] == 0).sum())
    else:
        raw_bits = phy.rx_receive(flat, n_packets, xor=False)[0]               # untimed, same entry point: raw decisions vs the transmitted (encoded) bits
    a = raw_bits[:, :nbytes].contiguous()
    b = tx_bits[:, :nbytes].contiguous()
    phy.ber_count(a, b, a.numel() * 8, cnt[:2])
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    errs, nb, sync_fail = (int(x) for x in cnt.cpu())
    del raw_bits, a, b

    # ---- end to end through the host-buffer API (pinned host in, packed bits back on the host)
    # default leg: uint8 PCM, the format of the reference's recordings (Final System Test.ipynb:85-86); the float32
    # and int16 legs are reported beside it.  Quantised legs are checked against the device path fed the same values.
    e2e_legs, e2e_keep = {}, {}
    e_steps = max(3, min(args.steps, 10))
    if not args.no_e2e:
        amax = float(data.abs().max())
        for key, dt, full_scale, off in (("e2e", torch.uint8, 120.0, 128.0), ("e2e_pcm16", torch.int16, 20000.0, 0.0), ("e2e_f32", torch.float32, None, 0.0)):
            try:
                if dt == torch.float32:
                    q = data
                else:
                    q = (torch.round(data * (full_scale / amax)) + off).to(dt)
                h_q = torch.empty(tuple(q.shape), dtype=dt).pin_memory()
                h_q.copy_(q)
                qf = q.to(torch.float32).contiguous() if dt != torch.float32 else None
                # chunks of 256 packets / streams; few long streams (a2-raw): halves, so that copies and kernels still overlap
                # while a chunk keeps enough streams for the detection-only matched filter (>= half the SM count)
                e_chunk = 256 if n_packets >= 1024 else max(96, (n_packets + 1) // 2)
                hr = HostReceiver(phy, n_packets, chunk=e_chunk, sample_dtype=dt, raw_T=T if raw else None)
                hr.run(h_q, xor=True)
                barrier()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record()
                for _ in range(e_steps):
                    resq = hr.run(h_q, xor=True)
                q1.record()
                barrier()
                q_ms = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=phy.device)
                if world > 1:
                    dist.all_reduce(q_ms, op=dist.ReduceOp.MAX)
                # the same entry points on the same samples already resident on the device
                if qf is None:
                    ref_dev = out_bits
                elif raw:
                    ref_dev = phy.receive_streams(q, 1, xor=True)["bits"]
                else:
                    ref_dev = phy.rx_receive_pcm(q.reshape(-1), n_packets, xor=True)[0]
                e2e_legs[key] = {"value": world * n_packets * phy.bits_per_packet * e_steps / (float(q_ms) * 1e-3) / 1e6, "unit": UNIT,
                                 "h2d_bytes_per_step": hr.h2d_bytes, "d2h_bytes_per_step": hr.d2h_bytes, "steps": e_steps,
                                 "api": "gf3b200.host.HostReceiver.run (%s %s in pinned host memory -> packed bits in pinned host memory, %d CUDA streams, chunks of %d)"
                                        % (str(dt).replace("torch.", ""), "raw streams" if raw else "packets", len(hr.streams), hr.chunk),
                                 "ingest": hr.ingest,
                                 "matches_device_result": bool(torch.equal(resq[:, :nbytes], ref_dev[:, :nbytes].cpu()))}
                if key == "e2e":
                    e2e_keep = {"bits": resq.clone(), "samples": qf}
                del hr, h_q, q
            except Exception as ex:   # report the failure instead of a made-up number
                e2e_legs[key] = {"value": None, "unit": UNIT, "error": repr(ex)}
        # the ceiling of any end-to-end number on this host: pinned-memory H2D bandwidth with every rank copying at once
        try:
            hb = torch.empty((256 << 20,), dtype=torch.uint8).pin_memory()
            db = torch.empty_like(hb, device=phy.device)
            db.copy_(hb, non_blocking=True)
            barrier()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(4):
                db.copy_(hb, non_blocking=True)
            h1.record()
            barrier()
            h_ms = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=phy.device)
            if world > 1:
                dist.all_reduce(h_ms, op=dist.ReduceOp.MAX)
            gbs = 4 * hb.numel() / (float(h_ms) * 1e-3) / 1e9
            for key, leg in e2e_legs.items():
                if leg.get("value"):
                    rate = leg["h2d_bytes_per_step"] * leg["steps"] / (n_packets * phy.bits_per_packet * leg["steps"] / (leg["value"] / world * 1e6)) / 1e9
                    leg["h2d_gbs_per_gpu"] = rate
                    leg["h2d_ceiling_gbs_per_gpu"] = gbs
                    leg["of_h2d_ceiling"] = rate / gbs
            del hb, db
        except Exception as ex:
            e2e_legs["e2e"]["h2d_probe_error"] = repr(ex)
    else:
        e2e_legs["e2e"] = {"value": None, "unit": UNIT, "error": "skipped (--no-e2e)"}
    sampler.stop()

    # ---- parity of the timed batch against the float64 oracle (rank 0's shard)
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            extra = []
            if e2e_keep:
                extra.append(("e2e", e2e_keep["bits"], e2e_keep["samples"]))
            parity = parity_check(phy, cfg, data, out_bits, raw, args.parity_packets if not raw else min(args.parity_packets, 16), extra)
        except Exception as ex:
            parity = {"error": repr(ex)}

    if rank == 0:
        bits_per_step = world * n_packets * phy.bits_per_packet
        syms_per_step = world * n_packets * (2 * phy.P + phy.L)
        sec = elapsed_ms * 1e-3
        peak, peak_src = peaks()
        demod_b, est_b, chain_b = alg_bytes(cfg)
        gb = data.numel() * 4 / 1e9
        prof = os.path.join(ROOT, "profiles", "ncu_demod_summary.json")
        traffic = None
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        if raw:
            xc_ms, pk_ms, rx_ms = seg[0], seg[1], seg[2]
            if not args.split_sync:      # one call: seg[0] is empty, seg[1] = matched filter + detection + offsets
                xc_ms, pk_ms = seg[1], 0.0
            Tb = 4.0 * T * n_packets                                        # one pass over the streams
            parts = -(-phy.chirp_len // 2048)
            one_kernel = parts <= 4          # longer chirps (the N = 4096 modes) take the multi-kernel matched filter (block spectra through a scratch array)
            dense = args.split_sync or args.dense_sync
            kb = 2 * Tb if dense else Tb        # all of P written (4T + 4T), or detection only: every sample read once (SURVEY 8d "ideal")
            if one_kernel:
                kname = "xcorr_fused_kernel (matched filter: 4T read + 4T written per stream when all of P is computed)"
            elif dense:
                kname = "xcorr_fwd_kernel + xcorr_mac_kernel + xcorr_acc_kernel (%d chirp partitions: block spectra and partition sums through scratch arrays, all of P written)" % parts
            else:
                kname = "xcorr_fwd_kernel (block spectra + group energies) + xcorr_bound_kernel + xcorr_acc_kernel on the blocks that can hold a candidate (%d chirp partitions)" % parts
            roof = {"bound": "hbm", "kernel": kname + ("" if args.split_sync else " + detection walk (%s)" % ("gf3_sync_streams" if args.dense_sync else "gf3_sync_detect: inverse transforms only where a candidate is possible")),
                    "achieved": kb / (xc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "traffic": traffic if dense else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kb, "avg_launch_ms": xc_ms,
                    "note": "the matched filter is bound by FP32 issue (144 flop per sample), not by HBM; " + ("4T read + 4T of P written" if dense else "detection only: 4T read, P written only where a candidate is possible (the dense form moves 8T: frac %.3f at this duration)" % (2 * Tb / (xc_ms * 1e-3) / 1e9 / peak)),
                    "stages_ms": {"matched_filter": xc_ms, "peak_pick_and_offsets": pk_ms, "receive_chain": rx_ms},
                    "sync": {"ms": xc_ms + pk_ms,
                             "ideal_4T": {"bytes": Tb, "frac": Tb / ((xc_ms + pk_ms) * 1e-3) / 1e9 / peak},
                             "two_pass_12T": {"bytes": 3 * Tb, "frac": 3 * Tb / ((xc_ms + pk_ms) * 1e-3) / 1e9 / peak}},
                    "receive_chain": {"achieved": chain_b * n_packets / (rx_ms * 1e-3) / 1e9, "frac": chain_b * n_packets / (rx_ms * 1e-3) / 1e9 / peak},
                    "whole_step": {"algorithmic_bytes": Tb + (phy.L * phy.Nd // 4 + 2 * phy.K * 8 + 4) * n_packets,
                                   "frac": (Tb + (phy.L * phy.Nd // 4 + 2 * phy.K * 8 + 4) * n_packets) * args.steps / sec / 1e9 / peak,
                                   "note": "every raw sample read once, bits + channel estimates written"}}
            roof["frac"] = roof["achieved"] / peak
        else:
            dem_ms = seg[2] + seg[3] if fused else seg[3]
            est_ms = seg[2]
            kernel_b = chain_b if fused else demod_b           # the fused launch moves the whole chain's bytes
            achieved = kernel_b * n_packets / (dem_ms * 1e-3) / 1e9
            roof = {"bound": "hbm",
                    "kernel": "rx_demod_kernel<FUSE_EST> (channel estimate + data symbols, one launch)" if fused else "rx_demod_kernel",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kernel_b * n_packets, "avg_launch_ms": dem_ms,
                    "estimate_kernel": None if fused else {"achieved": est_b * n_packets / (est_ms * 1e-3) / 1e9, "avg_launch_ms": est_ms,
                                                           "algorithmic_bytes_per_launch": est_b * n_packets},
                    "chain": {"achieved": chain_b * n_packets * args.steps / sec / 1e9,
                              "frac": chain_b * n_packets * args.steps / sec / 1e9 / peak,
                              "algorithmic_bytes_per_packet": chain_b}}
        line = {
            "metric": METRIC, "value": bits_per_step * args.steps / sec / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(desc, streams, world, gb),
            "check": {"bit_errors": errs, "bits_checked": nb, "ber": errs / max(nb, 1), "streams_sync_failed": sync_fail if raw else None,
                      "what": "timed path's decisions vs the transmitted bits, all ranks (NCCL sum)"},
            "symbols_per_s": syms_per_step * args.steps / sec,
            "roofline": roof, "parity": parity,
            "e2e": e2e_legs.get("e2e"), "e2e_pcm16": e2e_legs.get("e2e_pcm16"), "e2e_f32": e2e_legs.get("e2e_f32"),
            "gpu_launches": int(launches), "clocks": sampler.summary(), "numa": numa,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_thread(cfg, raw)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (default: the workload's)")
    ap.add_argument("--split-sync", action="store_true", help="c3-raw: time gf3_xcorr and gf3_peak_pick separately instead of gf3_sync_streams")
    ap.add_argument("--dense-sync", action="store_true", help="c3-raw: gf3_sync_streams (all of P computed) instead of gf3_sync_detect")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffer) legs (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed batch (profiling runs)")
    ap.add_argument("--parity-packets", type=int, default=64, help="packets of the timed batch that go through the oracle")
    args = ap.parse_args()
    cfg, streams, raw, desc = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    if args.impl == "reference":
        run_reference_arm(args, cfg, streams, raw, desc)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # not launched by torchrun: re-exec under it (one rank per GPU, NCCL)
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", str(port)] + sys.argv)
    run_gpu_arm(args, cfg, streams, raw, desc)


if __name__ == "__main__":
    main()
