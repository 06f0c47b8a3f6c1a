// gf3_rx.cu -- entry points of the receive chain (kernels: gf3_rx_kernels.cuh; the staged-input / PCM
// instantiations are compiled in gf3_rx_staged_*.cu).
#include "gf3_rx_kernels.cuh"

namespace gf3 {

// Staged input for float32 too?  Off unless asked for: GF3_RX_STAGED=1 (measured alternative to direct loads).
static bool staged_f32_default() {
    static const int v = [] { const char* e = getenv("GF3_RX_STAGED"); return e ? atoi(e) : 0; }();
    return v != 0;
}

// fmt / staged: sample format of `samples` (GF3_SAMPLE_*) and whether the symbols enter through the
// cp.async.bulk staging buffer (always for PCM; for float32 on request)
static int demod_common(const gf3_plan* plan, const void* samples, const int64_t* pkt_offset,
                        int64_t n_packets, const float* Hs, const float* He, const double* slope,
                        const uint8_t* xor2, uint8_t* bits, int64_t bits_stride, float* eq,
                        bool known_ch, int P_override, int L_override, void* stream,
                        const float* fuse_known = nullptr, int fmt = GF3_SAMPLE_F32, bool staged = false) {
    GF3_REQUIRE(plan && samples, "rx_demod: null plan or samples");
    GF3_REQUIRE(n_packets >= 0, "rx_demod: negative packet count");
    if (n_packets == 0) return GF3_OK;
    const gf3_params& p = plan->p;
    RxArgs a;
    memset(&a, 0, sizeof(a));
    a.samples = samples; a.pkt_offset = pkt_offset;
    a.Hs = reinterpret_cast<const float2*>(Hs); a.He = reinterpret_cast<const float2*>(He);
    a.slope = slope; a.xor2 = xor2; a.bits = bits; a.eq = reinterpret_cast<float2*>(eq);
    a.tw = plan->d_tw; a.bits_stride = bits_stride;
    a.cp = p.cp; a.lo = p.lo; a.hi = p.hi;
    a.P = P_override >= 0 ? P_override : p.n_pilots;
    a.L = L_override >= 0 ? L_override : p.packet_len;
    a.pkt_stride = (int64_t)(2 * a.P + a.L) * (p.N + p.cp);
    GF3_REQUIRE(a.L >= 1, "rx_demod: packet_len must be >= 1");
    GF3_REQUIRE(Hs != nullptr && (known_ch || (He != nullptr && slope != nullptr)), "rx_demod: null channel estimate");
    GF3_REQUIRE(bits || eq, "rx_demod: neither bits nor eq output requested");
    if (bits) {
        const int64_t need = (((int64_t)a.L * (p.hi - p.lo) * 2 + 31) / 32) * 4;
        GF3_REQUIRE(bits_stride % 4 == 0 && bits_stride >= need,
                    "rx_demod: bits_stride %lld must be a multiple of 4 and >= %lld", (long long)bits_stride, (long long)need);
        GF3_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 3) == 0, "rx_demod: bits_packed must be 4-byte aligned");
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (fuse_known) {
        a.known = reinterpret_cast<const float2*>(fuse_known);
        a.fit_lo = p.fit_lo; a.fit_hi = p.fit_hi;
    }
    if (fmt != GF3_SAMPLE_F32 || staged) {
        GF3_REQUIRE(!known_ch, "rx: the known-channel receiver takes float32 samples");
        if (fmt == GF3_SAMPLE_U8) return rx_staged_demod_u8(plan, a, n_packets, eq != nullptr, fuse_known != nullptr, st);
        if (fmt == GF3_SAMPLE_I16) return rx_staged_demod_i16(plan, a, n_packets, eq != nullptr, fuse_known != nullptr, st);
        GF3_REQUIRE(fmt == GF3_SAMPLE_F32, "rx: unknown sample format %d", fmt);
        return rx_staged_demod_f32(plan, a, n_packets, eq != nullptr, fuse_known != nullptr, st);
    }
    if (fuse_known) {           // estimate + data symbols in one launch (gf3_rx_receive)
        if (eq) { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, false, true, true>(plan, a, n_packets, st))); }
        else    { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, false, false, true>(plan, a, n_packets, st))); }
    }
    if (known_ch) {
        if (eq) { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, true, true>(plan, a, n_packets, st))); }
        else    { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, true, false>(plan, a, n_packets, st))); }
    } else {
        if (eq) { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, false, true>(plan, a, n_packets, st))); }
        else    { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P::LOGN, false, false>(plan, a, n_packets, st))); }
    }
    return GF3_OK;
}

}  // namespace gf3

using namespace gf3;

static int estimate_common(const gf3_plan* plan, const void* samples, int fmt, const int64_t* pkt_offset,
                           int64_t n_packets, const float* known, float* Hs, float* He,
                           double* slope, void* stream) {
    GF3_REQUIRE(plan && samples && known && Hs && He && slope, "rx_estimate: null argument");
    GF3_REQUIRE(n_packets >= 0 && n_packets <= 0x7fffffff, "rx_estimate: bad packet count");
    if (n_packets == 0) return GF3_OK;
    const gf3_params& p = plan->p;
    GF3_REQUIRE(p.n_pilots >= 1, "rx_estimate: n_pilots must be >= 1 (OFDM.py:424 short-circuits no_pilots == 0)");
    EstArgs a;
    a.samples = samples; a.pkt_offset = pkt_offset;
    a.known = reinterpret_cast<const float2*>(known);
    a.Hs = reinterpret_cast<float2*>(Hs); a.He = reinterpret_cast<float2*>(He); a.slope = slope;
    a.tw = plan->d_tw;
    a.pkt_stride = (int64_t)(2 * p.n_pilots + p.packet_len) * (p.N + p.cp);
    a.cp = p.cp; a.P = p.n_pilots; a.L = p.packet_len; a.fit_lo = p.fit_lo; a.fit_hi = p.fit_hi;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (fmt == GF3_SAMPLE_U8) return rx_estimate_u8(plan, a, n_packets, st);
    if (fmt == GF3_SAMPLE_I16) return rx_estimate_i16(plan, a, n_packets, st);
    GF3_REQUIRE(fmt == GF3_SAMPLE_F32, "rx_estimate: unknown sample format %d", fmt);
    GF3_DISPATCH_LOGN(plan->logN, return (launch_estimate<P>(plan, a, n_packets, st)));
    return GF3_OK;
}

extern "C" int gf3_rx_estimate(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                               int64_t n_packets, const float* known, float* Hs, float* He,
                               double* slope, void* stream) {
    return estimate_common(plan, samples, GF3_SAMPLE_F32, pkt_offset, n_packets, known, Hs, He, slope, stream);
}

extern "C" int gf3_rx_demod(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                            int64_t n_packets, const float* Hs, const float* He, const double* slope,
                            const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                            void* stream) {
    return demod_common(plan, samples, pkt_offset, n_packets, Hs, He, slope, xor2, bits_packed,
                        bits_stride, eq, false, -1, -1, stream);
}

template <class P0>
static int receive_is_fused(const gf3_plan* plan) {
    using C = DemodCfg<P0::LOGN>;
    using P = typename C::Plan;
    constexpr int NT = (P::T > C::NT) ? P::T : C::NT, SF = NT / P::T;
    if (demod_est_par<P, NT>() < 2 && !GF3_FUSE_SEQUENTIAL) return 0;
    if (P::LOGN == 12 && !GF3_FUSE12 && !GF3_FUSE_SEQUENTIAL) return 0;
    const gf3_params& p = plan->p;
    const int K = P::M - 1, Nd = p.hi - p.lo;
    const int flo = p.fit_lo < 0 ? 0 : (p.fit_lo > K ? K : p.fit_lo), fhi = p.fit_hi < flo ? flo : (p.fit_hi > K ? K : p.fit_hi);
    int flush = SF;
    while ((flush * Nd) % 16 != 0 || flush < 8) flush += SF;     // as in launch_demod
    return (size_t)2 * (fhi - flo) * sizeof(double) <= ((((size_t)flush * Nd + 15) & ~(size_t)15) + 16) ? 1 : 0;
}

extern "C" int gf3_rx_receive_is_fused(const gf3_plan* plan) {
    if (!plan || plan->p.n_pilots < 1) return 0;
    GF3_DISPATCH_LOGN(plan->logN, return receive_is_fused<P>(plan));
    return 0;
}

static int receive_common(const gf3_plan* plan, const void* samples, int fmt, bool staged, const int64_t* pkt_offset,
                          int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                          const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq, void* stream) {
    GF3_REQUIRE(plan && known && Hs && He && slope, "rx_receive: null argument");
    GF3_REQUIRE(plan->p.n_pilots >= 1, "rx_receive: n_pilots must be >= 1 (OFDM.py:424 short-circuits no_pilots == 0)");
    GF3_REQUIRE(fmt == GF3_SAMPLE_F32 || fmt == GF3_SAMPLE_I16 || fmt == GF3_SAMPLE_U8, "rx_receive: unknown sample format %d", fmt);
    int rc = demod_common(plan, samples, pkt_offset, n_packets, Hs, He, slope, xor2, bits_packed, bits_stride, eq,
                          false, -1, -1, stream, known, fmt, staged);
    if (rc != GF3_FUSE_UNFIT) return rc;
    // geometry the fused kernel cannot hold (N = 4096, or a very wide fit window on a short flush period): two launches
    rc = estimate_common(plan, samples, fmt, pkt_offset, n_packets, known, Hs, He, slope, stream);
    if (rc) return rc;
    return demod_common(plan, samples, pkt_offset, n_packets, Hs, He, slope, xor2, bits_packed, bits_stride, eq,
                        false, -1, -1, stream, nullptr, fmt, staged);
}

extern "C" int gf3_rx_receive(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                              int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                              const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                              void* stream) {
    return receive_common(plan, samples, GF3_SAMPLE_F32, staged_f32_default(), pkt_offset, n_packets, known, Hs, He, slope, xor2,
                          bits_packed, bits_stride, eq, stream);
}

extern "C" int gf3_rx_receive_pcm(const gf3_plan* plan, const void* samples, int32_t sample_format, const int64_t* pkt_offset,
                                  int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                                  const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                                  void* stream) {
    GF3_REQUIRE(samples != nullptr, "rx_receive_pcm: null samples");
    return receive_common(plan, samples, sample_format, true, pkt_offset, n_packets, known, Hs, He, slope, xor2,
                          bits_packed, bits_stride, eq, stream);
}

extern "C" int gf3_rx_known_channel(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                                    int64_t n_packets, const float* Hinv, const uint8_t* xor2,
                                    uint8_t* bits_packed, int64_t bits_stride, float* eq, void* stream) {
    return demod_common(plan, samples, pkt_offset, n_packets, Hinv, nullptr, nullptr, xor2,
                        bits_packed, bits_stride, eq, true, -1, -1, stream);
}

extern "C" int gf3_rx_spectrum(const gf3_plan* plan, const float* samples, const int64_t* sym_offset,
                               int64_t n_symbols, float* out, void* stream) {
    GF3_REQUIRE(plan && out, "rx_spectrum: null argument");
    // every symbol is a one-symbol "packet" with a unit channel: eq output == FFT bins 1..K
    return demod_common(plan, samples, sym_offset, n_symbols, reinterpret_cast<const float*>(plan->d_ones),
                        nullptr, nullptr, nullptr, nullptr, 0, out, true, 0, 1, stream);
}
