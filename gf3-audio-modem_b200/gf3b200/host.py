"""Host-buffer entry points: the calls a user with samples in HOST memory makes.  Host->device
copies, the fused kernels and the device->host read of the result are pipelined over CUDA streams
in chunks of packets, so PCIe transfer overlaps compute."""
import torch


class HostReceiver:
    """Receive chain for packets held in pinned host memory (float32, [n_packets, pkt_samples]).

    out = HostReceiver(phy, n_packets, chunk).run(sym_host) -> packed bits, pinned uint8
    [n_packets, bits_stride].  Every call moves n_packets*pkt_samples*4 bytes H2D and
    n_packets*bits_stride bytes D2H.
    """

    def __init__(self, phy, n_packets, chunk=256, n_streams=3, sample_dtype=torch.float32):
        """sample_dtype: torch.float32, or torch.int16 / torch.uint8 for PCM as recorded (converted
        on the device by gf3_pcm_to_f32, so 2 or 1 bytes per sample cross PCIe instead of 4)."""
        self.phy, self.n_packets, self.chunk = phy, n_packets, min(chunk, n_packets)
        self.sample_dtype = sample_dtype
        self.streams = [torch.cuda.Stream(device=phy.device) for _ in range(n_streams)]
        self.d_in = [torch.empty((self.chunk, phy.pkt_samples), dtype=torch.float32, device=phy.device) for _ in self.streams]
        self.d_pcm = None
        if sample_dtype != torch.float32:
            self.d_pcm = [torch.empty((self.chunk, phy.pkt_samples), dtype=sample_dtype, device=phy.device) for _ in self.streams]
        self.d_out = [torch.empty((self.chunk, phy.bits_stride), dtype=torch.uint8, device=phy.device) for _ in self.streams]
        self.h_out = torch.empty((n_packets, phy.bits_stride), dtype=torch.uint8).pin_memory()
        self._done = torch.cuda.Event()
        self.h2d_bytes = n_packets * phy.pkt_samples * torch.empty((), dtype=sample_dtype).element_size()
        self.d2h_bytes = n_packets * phy.bits_stride

    def run(self, sym_host, xor=True):
        phy = self.phy
        assert sym_host.is_pinned() and sym_host.dtype == self.sample_dtype and sym_host.shape == (self.n_packets, phy.pkt_samples)
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)
        for i, p0 in enumerate(range(0, self.n_packets, self.chunk)):
            n = min(self.chunk, self.n_packets - p0)
            k = i % len(self.streams)
            with torch.cuda.stream(self.streams[k]):
                d = self.d_in[k][:n]
                if self.d_pcm is None:
                    d.copy_(sym_host[p0:p0 + n], non_blocking=True)
                else:
                    self.d_pcm[k][:n].copy_(sym_host[p0:p0 + n], non_blocking=True)
                    phy.pcm_to_f32(self.d_pcm[k][:n], out=d)
                phy.rx_receive(d.reshape(-1), n, xor=xor, out=self.d_out[k][:n])        # estimate + data symbols, one launch
                self.h_out[p0:p0 + n].copy_(self.d_out[k][:n], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        # the result is returned as finished host memory: wait for the last device-to-host copy
        self._done.record(cur)
        self._done.synchronize()
        return self.h_out
