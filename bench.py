#!/usr/bin/env python
"""bench.py -- headline benchmark of the GF3 B200 physical layer (BASELINE.json metric:
demodulated Mbit/s and OFDM symbols/s, % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|a2] [--impl reference]

A "step" is one pass of the receive chain (gf3_rx_receive: channel estimate + data symbols in ONE
kernel launch for N <= 2048, gf3_rx_estimate + gf3_rx_demod for N = 4096) over one batch of
synthetic received packets.  Default workload = BASELINE.json
configs[2] ("C3": 4096 independent streams x 1 packet, N=1024, CP=32, 511 data bins, P=20, L=180,
random 30-tap multipath + AWGN 20 dB) -- the configuration the metric is quoted on; configs[0]/[1]
are single-stream decodes and are covered as parity tests.  One process per GPU (torchrun for
N > 1): streams are sharded across ranks with no data-path collective (weak scaling); NCCL
all-reduces only the bit-error counters.  Prints ONE JSON line on rank 0.
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

WORKLOADS = {
    # name: (params, streams per GPU, description)
    # fit window: the reference hard-codes carrier indices [500:1000) for K=2047 (OFDM.py:462); at
    # K=511 that clips to 11 edge bins and the drift estimate is noise, so the N=1024 workload uses
    # the same band fraction, [125:250) ("BER-quality sweeps may expose fit_lo/fit_hi")
    "c3": (dict(N=1024, cp=32, lo=1, hi=512, n_pilots=20, packet_len=180, fit_lo=125, fit_hi=250), 4096,
           "C3 (BASELINE.json configs[2]): 4096 streams x 1 packet, N=1024, CP=32, Nd=511, P=20, L=180, fit window [125:250), random 30-tap multipath + AWGN 20 dB"),
    "c4": (dict(N=4096, cp=704, lo=1, hi=2047, n_pilots=20, packet_len=180), 512,
           "C4 (BASELINE.json configs[3], mode B1): 512 streams x 1 packet, N=4096, CP=704, Nd=2046, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
    "w2048": (dict(N=2048, cp=64, lo=1, hi=1024, n_pilots=20, packet_len=180, fit_lo=250, fit_hi=500), 2048,
              "W2048 (parity-test geometry): 2048 streams x 1 packet, N=2048, CP=64, Nd=1023, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
    "a2": (dict(N=4096, cp=224, lo=100, hi=1500, n_pilots=20, packet_len=180), 512,
           "A2 (mode of the real recording): 512 streams x 1 packet, N=4096, CP=224, Nd=1400, P=20, L=180, random 30-tap multipath + AWGN 20 dB"),
}
METRIC, UNIT = "demodulated_mbit_per_s", "Mbit/s"


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


# ------------------------------------------------------------------------------ CPU arms
def _cpu_packets(cfg, n, seed=99):
    """Synthetic received packets of the workload, made on the CPU with the oracle's transmit
    chain + scipy FIR + AWGN (float64 [n, 2P+L, N+cp])."""
    import numpy as np
    from scipy.signal import lfilter
    from oracle import gf3_oracle as orc
    import gf3b200.phy as _phy
    ks = _phy.default_known_sequence()
    p = orc.Params(N=cfg["N"], cp=cfg["cp"], lo=cfg["lo"], hi=cfg["hi"], n_pilots=cfg["n_pilots"],
                   packet_len=cfg["packet_len"], known_sequence=ks, encoding="XOR",
                   fit_lo=cfg.get("fit_lo", 500), fit_hi=cfg.get("fit_hi", 1000))
    rng = np.random.default_rng(seed)
    out = np.empty((n, p.syms_per_packet, p.sym_len))
    bits_all = []
    k = np.arange(30)
    for i in range(n):
        bits = rng.integers(0, 2, p.data_bits_per_symbol * p.packet_len)
        tx = orc.transmit(p, bits, rng=_RngShim(rng))
        t = rng.normal(0.0, np.sqrt(np.exp(-k / 5.0)))
        t[0] = abs(t[0]) + 1.0
        t /= np.sqrt(np.sum(t * t))
        y = lfilter(t, 1.0, tx)
        pkt = y[p.chirp_length:p.chirp_length + p.packet_samples]
        pkt = pkt + rng.normal(0, np.sqrt(np.mean(pkt ** 2)) * 0.1, len(pkt))       # 20 dB
        out[i] = pkt.reshape(p.syms_per_packet, p.sym_len)
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


_CPU_JOB = None      # (params, packets) inherited by the forked workers: nothing is pickled per step


def _cpu_worker(_):
    from oracle import gf3_oracle as orc
    p, rx = _CPU_JOB
    return len(orc.receive_symbols(p, rx)["bits"])


def cpu_baseline_single_thread(cfg, budget_s=12.0):
    """The numpy oracle (float64 port of the reference's receive chain) on one host thread over a
    bounded sample of the workload."""
    import numpy as np
    from oracle import gf3_oracle as orc
    p, rx, bits = _cpu_packets(cfg, 16)
    out = orc.receive_symbols(p, rx[:2])                       # warm-up
    t0 = time.perf_counter()
    nbits = npk = 0
    while time.perf_counter() - t0 < budget_s:
        out = orc.receive_symbols(p, rx)
        nbits += len(out["bits"])
        npk += rx.shape[0]
    dt = time.perf_counter() - t0
    errs = int(np.sum(out["bits"] != np.concatenate(bits)))
    return dict(value=nbits / dt / 1e6, unit=UNIT, cores=1, kind="port",
                sample="%d packets (%d OFDM symbols) of the workload in %.1f s, float64 numpy oracle (oracle/gf3_oracle.py), 1 thread; %d bit errors in the last 16 packets"
                       % (npk, npk * p.syms_per_packet, dt, errs),
                symbols_per_s=npk * p.syms_per_packet / dt)


def run_reference_arm(args, cfg, desc):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself is
    Python under /root/reference and cannot travel to the GPU box, so this times the numpy oracle
    port (validated <= 1e-12 against the reference, tests/test_oracle_golden.py) on all host cores,
    one process per core, each step a bounded sample of the workload."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 16
    global _CPU_JOB
    p, rx, _ = _cpu_packets(cfg, per_core)
    _CPU_JOB = (p, rx)
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample_per_step": "%d packets (%d per core x %d cores)" % (cores * per_core, per_core, cores)},
        "symbols_per_s": npk * p.syms_per_packet / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d packets per step x %d steps, one process per host core, float64 numpy oracle port of OFDM.py:391-609" % (cores * per_core, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
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


# ------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, cfg, streams, desc):
    import torch
    import torch.distributed as dist
    import gf3b200
    from gf3b200 import synth
    from gf3b200.host import HostReceiver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    phy = gf3b200.Phy(**cfg)
    # ---- synthetic input, generated on the device by the package's own tx + channel kernels
    gen_chunk = 512
    sym = torch.empty((streams, phy.pkt_samples), dtype=torch.float32, device=phy.device)
    tx_bits = torch.empty((streams, phy.bits_stride), dtype=torch.uint8, device=phy.device)
    for s0 in range(0, streams, gen_chunk):
        n = min(gen_chunk, streams - s0)
        b = synth.make_batch(phy, n, 1, snr_db=20.0, seed=1234, first_stream=rank * streams + s0)
        sym[s0:s0 + n] = synth.packets_from_streams(phy, b)
        tx_bits[s0:s0 + n] = b["bits"][:, 0]
        del b
    torch.cuda.synchronize()
    n_packets = streams
    out_bits = torch.empty((n_packets, phy.bits_stride), dtype=torch.uint8, device=phy.device)
    flat = sym.reshape(-1)

    fused = phy.fused_receive       # one launch for the whole chain (N <= 2048), else estimate + demod

    def step(events=None):
        if events is not None:
            events[0].record()
        if fused:
            if events is not None:
                events[1].record()
            phy.rx_receive(flat, n_packets, xor=True, out=out_bits)             # XOR decode fused (Final System Test uses encoding="XOR")
        else:
            Hs, He, slope = phy.rx_estimate(flat, n_packets)
            if events is not None:
                events[1].record()
            phy.rx_demod(flat, n_packets, Hs, He, slope, xor=True, out=out_bits)
        if events is not None:
            events[2].record()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
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
    est_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    dem_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([elapsed_ms, dem_ms, est_ms], dtype=torch.float64, device=phy.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, dem_ms_max, est_ms_max = (float(x) for x in t.cpu())

    # ---- correctness of what was timed: BER against the transmitted bits (NCCL sum of counters)
    cnt = torch.zeros(2, dtype=torch.int64, device=phy.device)
    nbytes = (phy.bits_per_packet + 7) // 8
    raw_bits = phy.rx_receive(flat, n_packets, xor=False)[0]               # untimed, same entry point: raw decisions vs the transmitted (encoded) bits
    a = raw_bits[:, :nbytes].contiguous()
    b = tx_bits[:, :nbytes].contiguous()
    phy.ber_count(a, b, a.numel() * 8, cnt)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    errs, nb = (int(x) for x in cnt.cpu())

    # ---- end to end through the host-buffer API (pinned host in, packed bits back on the host)
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        hr = HostReceiver(phy, n_packets, chunk=256)
        h_sym = torch.empty((n_packets, phy.pkt_samples), dtype=torch.float32).pin_memory()
        h_sym.copy_(sym)
        torch.cuda.synchronize()
        hr.run(h_sym, xor=True)
        barrier()
        e_steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e_steps):
            res = hr.run(h_sym, xor=True)
        e1.record()
        barrier()
        e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=phy.device)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(res[:, :nbytes], out_bits[:, :nbytes].cpu()))
        e2e = {"value": world * n_packets * phy.bits_per_packet * e_steps / (float(e_ms) * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": hr.h2d_bytes, "d2h_bytes_per_step": hr.d2h_bytes, "steps": e_steps,
               "api": "gf3b200.host.HostReceiver.run (pinned float32 host packets -> packed bits in pinned host memory, 3 CUDA streams)",
               "matches_device_result": same}
    except Exception as ex:   # report the failure instead of a made-up number
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)}
    # ---- the same path fed with 16-bit PCM host buffers (the native format of the recordings):
    # half / a quarter of the PCIe bytes; reported separately because the samples are quantised (the
    # reference's own recordings are 8-bit PCM wav files, Final System Test.ipynb:85-86)
    e2e_pcm = {}
    if e2e.get("value") and not args.no_e2e:
        del hr, h_sym
        for key, dt, full_scale, off in (("e2e_pcm16", torch.int16, 20000.0, 0.0), ("e2e_pcm8", torch.uint8, 120.0, 128.0)):
            try:
                scale = full_scale / float(sym.abs().max())
                sym_q = (torch.round(sym * scale) + off).to(dt)
                h_q = torch.empty(sym_q.shape, dtype=dt).pin_memory()
                h_q.copy_(sym_q)
                qf = sym_q.to(torch.float32).reshape(-1)
                ref_q = phy.rx_receive(qf, n_packets, xor=True)[0]
                hq = HostReceiver(phy, n_packets, chunk=256, sample_dtype=dt)
                hq.run(h_q, xor=True)
                barrier()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record()
                for _ in range(e_steps):
                    resq = hq.run(h_q, xor=True)
                q1.record()
                barrier()
                q_ms = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=phy.device)
                if world > 1:
                    dist.all_reduce(q_ms, op=dist.ReduceOp.MAX)
                e2e_pcm[key] = {"value": world * n_packets * phy.bits_per_packet * e_steps / (float(q_ms) * 1e-3) / 1e6, "unit": UNIT,
                                "h2d_bytes_per_step": hq.h2d_bytes, "d2h_bytes_per_step": hq.d2h_bytes, "steps": e_steps,
                                "note": "%s PCM host buffers, converted on the device (gf3_pcm_to_f32)" % str(dt).replace("torch.", ""),
                                "matches_device_result": bool(torch.equal(resq[:, :nbytes], ref_q[:, :nbytes].cpu()))}
                del sym_q, qf, ref_q, hq, h_q
            except Exception as ex:
                e2e_pcm[key] = {"value": None, "error": repr(ex)}
    sampler.stop()

    if rank == 0:
        bits_per_step = world * n_packets * phy.bits_per_packet
        syms_per_step = world * n_packets * (2 * phy.P + phy.L)
        sec = elapsed_ms * 1e-3
        peak, peak_src = peaks()
        demod_b, est_b, chain_b = alg_bytes(cfg)
        kernel_b = chain_b if fused else demod_b           # the fused launch moves the whole chain's bytes
        achieved = kernel_b * n_packets / (dem_ms_max * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_demod_summary.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": bits_per_step * args.steps / sec / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "streams_per_gpu": streams, "packets_per_step": world * n_packets,
                       "l2_policy": "input batch (%.2f GB per GPU) is larger than the 126 MB L2; no flush needed" % (sym.numel() * 4 / 1e9),
                       "bit_errors": errs, "bits_checked": nb, "ber": errs / max(nb, 1)},
            "symbols_per_s": syms_per_step * args.steps / sec,
            "roofline": {"bound": "hbm",
                         "kernel": "rx_demod_kernel<FUSE_EST> (channel estimate + data symbols, one launch)" if fused else "rx_demod_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kernel_b * n_packets, "avg_launch_ms": dem_ms_max,
                         "estimate_kernel": None if fused else {"achieved": est_b * n_packets / (est_ms_max * 1e-3) / 1e9, "avg_launch_ms": est_ms_max,
                                                                "algorithmic_bytes_per_launch": est_b * n_packets},
                         "chain": {"achieved": chain_b * n_packets * world * args.steps / sec / 1e9 / world,
                                   "frac": chain_b * n_packets * args.steps / sec / 1e9 / peak,
                                   "algorithmic_bytes_per_packet": chain_b}},
            "e2e": e2e, "e2e_pcm16": e2e_pcm.get("e2e_pcm16"), "e2e_pcm8": e2e_pcm.get("e2e_pcm8"), "gpu_launches": int(launches), "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_thread(cfg)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffer) leg (profiling runs)")
    args = ap.parse_args()
    cfg, streams, desc = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    if args.impl == "reference":
        run_reference_arm(args, cfg, desc)
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
    run_gpu_arm(args, cfg, streams, desc)


if __name__ == "__main__":
    main()
