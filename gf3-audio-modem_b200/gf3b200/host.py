"""Host-buffer entry points: the calls a user with samples in HOST memory makes.  Host->device
copies, the fused kernels and the device->host read of the result are pipelined over CUDA streams
in chunks of packets (or raw streams), so PCIe transfer overlaps compute."""
import ctypes
import os

import torch


def bind_to_gpu_numa(gpu_index):
    """Place this process next to its GPU before it allocates pinned host buffers: CPU affinity and the
    memory policy of later allocations go to the NUMA node the GPU's PCIe root hangs off (sysfs
    numa_node of the device's bus id).  On a single-node (or unreported) topology this is a no-op.
    Returns a dict describing what was found and done (bench.py prints it)."""
    info = {"gpu": gpu_index, "node": None, "nodes_online": None, "cpus_bound": None, "mempolicy": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        info["bus"] = bus
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        info["node"] = node
        with open("/sys/devices/system/node/online") as f:
            info["nodes_online"] = f.read().strip()
        if node < 0 or info["nodes_online"] in ("0", ""):
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = len(allowed)
        # set_mempolicy(MPOL_PREFERRED, {node}): later allocations (the pinned buffers) come from this node
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
    except Exception as ex:     # topology not visible from inside the container: say so, change nothing
        info["error"] = repr(ex)
    return info


class HostReceiver:
    """Receive chain for samples held in pinned host memory.

    packets:     sym_host [n_packets, pkt_samples]            (raw_T = None)
    raw streams: r_host   [n_streams, raw_T], one packet each (raw_T = samples per stream): matched filter,
                 detection rule and packet offsets run on the device too (Phy.receive_streams)
    sample_dtype float32, or int16 / uint8 PCM as recorded (Final System Test.ipynb:85-86 reads an 8-bit wav):
    2 or 1 bytes per sample cross PCIe instead of 4.

    out = HostReceiver(...).run(host_samples) -> packed bits, pinned uint8 [n, bits_stride], complete when
    run() returns.  Every call moves n * samples * itemsize bytes H2D and n * bits_stride bytes D2H.
    """

    def __init__(self, phy, n_packets, chunk=256, n_streams=3, sample_dtype=torch.float32, raw_T=None):
        self.phy, self.n_packets, self.chunk = phy, n_packets, min(chunk, n_packets)
        self.sample_dtype, self.raw_T = sample_dtype, raw_T
        self.row = raw_T if raw_T is not None else phy.pkt_samples
        self.streams = [torch.cuda.Stream(device=phy.device) for _ in range(n_streams)]
        pcm = sample_dtype != torch.float32
        # no float copy of a PCM batch ever exists in HBM
        self.in_kernel = pcm                     # PCM is converted inside the kernels (matched filter and receive chain alike)
        self.ingest = "float32" if not pcm else "in-kernel (gf3_sync_streams / gf3_rx_receive_pcm read the PCM samples)"
        self.d_in = None if self.in_kernel else [torch.empty((self.chunk, self.row), dtype=torch.float32, device=phy.device) for _ in self.streams]
        self.d_pcm = [torch.empty((self.chunk, self.row), dtype=sample_dtype, device=phy.device) for _ in self.streams] if pcm else None
        self.d_out = [torch.empty((self.chunk, phy.bits_stride), dtype=torch.uint8, device=phy.device) for _ in self.streams]
        self.h_out = torch.empty((n_packets, phy.bits_stride), dtype=torch.uint8).pin_memory()
        self._done = torch.cuda.Event()
        self.h2d_bytes = n_packets * self.row * torch.empty((), dtype=sample_dtype).element_size()
        self.d2h_bytes = n_packets * phy.bits_stride

    def run(self, host, xor=True):
        phy = self.phy
        assert host.is_pinned() and host.dtype == self.sample_dtype and host.shape == (self.n_packets, self.row)
        cur = torch.cuda.current_stream(phy.device)
        for s in self.streams:
            s.wait_stream(cur)
        for i, p0 in enumerate(range(0, self.n_packets, self.chunk)):
            n = min(self.chunk, self.n_packets - p0)
            k = i % len(self.streams)
            with torch.cuda.stream(self.streams[k]):
                if self.d_pcm is None:
                    d = self.d_in[k][:n]
                    d.copy_(host[p0:p0 + n], non_blocking=True)
                else:
                    self.d_pcm[k][:n].copy_(host[p0:p0 + n], non_blocking=True)
                    if not self.in_kernel:
                        d = self.d_in[k][:n]
                        phy.pcm_to_f32(self.d_pcm[k][:n], out=d)
                if self.raw_T is not None:
                    phy.receive_streams(self.d_pcm[k][:n] if self.in_kernel else d, 1, xor=xor, out=self.d_out[k][:n])
                elif self.in_kernel:
                    phy.rx_receive_pcm(self.d_pcm[k][:n].reshape(-1), n, xor=xor, out=self.d_out[k][:n])
                else:
                    phy.rx_receive(d.reshape(-1), n, xor=xor, out=self.d_out[k][:n])        # estimate + data symbols, one launch
                self.h_out[p0:p0 + n].copy_(self.d_out[k][:n], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        # the result is returned as finished host memory: wait for the last device-to-host copy
        self._done.record(cur)
        self._done.synchronize()
        return self.h_out
