// gf3_tx.cu -- the transmit chain as fused sm_100a kernels.
//
//   tx_symbols_kernel : bits -> Gray QPSK -> Hermitian spectrum -> inverse real FFT -> CP -> gain,
//                       written straight into the packet frame    (OFDM.py:191-226, 322-323, 256)
//   tx_frame_kernel   : chirp preamble(s) and the 2P known symbols of every packet (OFDM.py:244-259)
//
// The inverse real FFT of N samples runs as one forward M = N/2 point complex FFT on conjugated
// input (ifft(Z) = conj(fft(conj Z))/M) in the same register-resident engine as the receiver.
#include "gf3_common.cuh"
#include "gf3_fft.cuh"

namespace gf3 {

constexpr int kTxThreads = 256;

struct TxArgs {
    const uint8_t* bits;        // [n_streams, pk_per_stream, bits_stride]   (null for the known symbol)
    const float2* filler;       // [n_streams, K - Nd]
    const float2* known;        // [K] (used when bits == null)
    const float2* tw;
    float* out;                 // [n_streams, out_stride]
    int64_t bits_stride, out_stride;
    int64_t pk_per_stream;
    int cp, lo, hi, P, L, chirp_len;
    float gain;                 // tx_gain / N
    int batches_per_packet;     // ceil(L / SF)
};

// QPSK point of the encoded bit pair of data carrier c in symbol l (OFDM.py:72-77):
// (b0,b1) -> ((1-2 b1) + j (1-2 b0)) / sqrt(2)
__device__ __forceinline__ float2 qpsk_from_bits(const uint8_t* __restrict__ sbits, int bitpos) {
    const unsigned byte = sbits[bitpos >> 3];
    const int sh = 6 - (bitpos & 7);                   // bitpos is even: b0 at 7-(g&7), b1 one below
    const unsigned b0 = (byte >> (sh + 1)) & 1u, b1 = (byte >> sh) & 1u;
    const float h = 0.70710678118654752440f;
    return make_float2(b1 ? -h : h, b0 ? -h : h);
}

// One CTA = SF symbols of one packet.  Phase B' (thread <-> bin pair) builds conj(Z) in smem,
// phase A' runs the FFT, the epilogue writes x[2m] = Re Y[m], x[2m+1] = -Im Y[m] (times gain)
// plus the cyclic prefix.
template <class P, bool KNOWN_SYMBOL>
__global__ void __launch_bounds__(kTxThreads, 2) tx_symbols_kernel(const TxArgs a) {
    constexpr int NT = kTxThreads, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, K = M - 1;
    constexpr int SF = NT / T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + SF * MP;
    uint8_t* sbits = reinterpret_cast<uint8_t*>(tw + P::TW_TOTAL);     // [SF][bytes_per_sym + 2]

    const int tid = threadIdx.x;
    const int Nd = a.hi - a.lo;
    const int symlen = N + a.cp;
    int64_t pktg = 0;           // global packet index (stream * pk_per_stream + packet)
    int l_first = 0, nsym = 1;
    if constexpr (!KNOWN_SYMBOL) {
        pktg = blockIdx.x / a.batches_per_packet;
        l_first = (blockIdx.x % a.batches_per_packet) * SF;
        nsym = min(SF, a.L - l_first);
    }
    const int64_t stream = pktg / a.pk_per_stream, pk = pktg % a.pk_per_stream;

    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // stage the packed bits of the nsym symbols (bit offset l*2Nd is not byte aligned in general)
    const int sym_bytes = (2 * Nd + 7) / 8 + 1;
    if constexpr (!KNOWN_SYMBOL) {
        const uint8_t* pb = a.bits + pktg * a.bits_stride;
        for (int i = tid; i < nsym * sym_bytes; i += NT) {
            const int s = i / sym_bytes, o = i % sym_bytes;
            const int64_t byte0 = ((int64_t)(l_first + s) * 2 * Nd) >> 3;
            const int64_t idx = byte0 + o;
            sbits[s * sym_bytes + o] = idx < a.bits_stride ? pb[idx] : 0;
        }
    }
    __syncthreads();

    // ---- phase B': Hermitian spectrum -> conj(Z[k]) for the packed inverse real FFT
    //   E = X[k] + conj X[M-k],  O = (X[k] - conj X[M-k]) e^{+2 pi i k/N},  Z = E + jO
    for (int item = tid; item < SF * (M / 2 + 1); item += NT) {
        const int s = item / (M / 2 + 1), k = item % (M / 2 + 1), km = M - k;
        float2 X1 = make_float2(0.f, 0.f), X2 = make_float2(0.f, 0.f);
        if (s < nsym) {
            auto bin = [&](int kk) -> float2 {
                if (kk < 1 || kk > K) return make_float2(0.f, 0.f);              // DC and Nyquist stay 0 (OFDM.py:209)
                if constexpr (KNOWN_SYMBOL) return a.known[kk - 1];
                else {
                    if (kk >= a.lo && kk < a.hi) {
                        const int g = (int)((((int64_t)(l_first + s) * 2 * Nd) & 7) + 2 * (kk - a.lo));
                        return qpsk_from_bits(sbits + s * sym_bytes, g);
                    }
                    // np.delete(carriers, data_carriers-1) keeps ascending order (OFDM.py:49,213)
                    const int u = kk < a.lo ? kk - 1 : kk - 1 - Nd;
                    return a.filler[stream * (K - Nd) + u];
                }
            };
            X1 = bin(k);
            X2 = bin(km);
        }
        float sn, cs;
        sincospif(2.0f * (float)k / (float)N, &sn, &cs);                          // e^{+j theta}
        const float2 E = make_float2(X1.x + X2.x, X1.y - X2.y);
        const float2 D = make_float2(X1.x - X2.x, X1.y + X2.y);
        const float2 O = cmul(D, make_float2(cs, sn));
        // Z[k] = E + jO ;  Z[M-k] = conj(E) + j conj(O)... derived from the same pair:
        const float2 Zk = make_float2(E.x - O.y, E.y + O.x);
        const float2 Zm = make_float2(E.x + O.y, O.x - E.y);                       // conj(E - jO) = conj(E) + j conj(O)
        float2* zs = zbuf + s * MP;
        if (k < M) zs[zpad<P>(k)] = cconj(Zk);
        if (k != 0 && km != k) zs[zpad<P>(km)] = cconj(Zm);
        else if (k == 0) { /* Z[M] aliases Z[0]; nothing to store */ }
    }
    __syncthreads();

    // ---- phase A': forward FFT of conj(Z)
    {
        const int g = tid / T, t = tid % T;
        float2 x[R];
        float2* zs = zbuf + g * MP;
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = zs[zpad<P>(t + i * T)];
        group_sync<P, NT>(g);
        fft_forward<P, NT>(x, zs, tw, t, g);
    }
    __syncthreads();

    // ---- epilogue: time samples with cyclic prefix (OFDM.py:221-226), gain (OFDM.py:256)
    for (int item = tid; item < nsym * M; item += NT) {
        const int s = item / M, m = item % M;
        const float2 y = zbuf[s * MP + zpad<P>(m)];
        const float2 v = make_float2(y.x * a.gain, -y.y * a.gain);
        float* o;
        if constexpr (KNOWN_SYMBOL) o = a.out;
        else o = a.out + stream * a.out_stride + pk * ((int64_t)a.chirp_len + (int64_t)(2 * a.P + a.L) * symlen)
                 + a.chirp_len + (int64_t)(a.P + l_first + s) * symlen;
        const int n0 = 2 * m - (N - a.cp);                       // position of this pair inside the cyclic prefix
        if (((reinterpret_cast<uintptr_t>(o) | (uintptr_t)(a.cp * 4)) & 7) == 0) {       // 8-byte aligned symbol and even CP
            *reinterpret_cast<float2*>(o + a.cp + 2 * m) = v;
            if (n0 >= 0) *reinterpret_cast<float2*>(o + n0) = v;
        } else {
            o[a.cp + 2 * m] = v.x;
            o[a.cp + 2 * m + 1] = v.y;
            if (n0 >= 0) o[n0] = v.x;
            if (n0 + 1 >= 0) o[n0 + 1] = v.y;
        }
    }
}

// Chirp preambles and known symbols: plain copies.
struct FrameArgs {
    const float* chirp;         // [chirp_len]
    const float* known_time;    // [N + cp], gain applied
    float* out;
    int64_t out_stride, pk_per_stream, n_streams;
    int symlen, P, L, chirp_len, gx;
};

// row = blockIdx.x / gx = stream * (pk_per_stream + 1) + packet; the extra "packet" is the trailing chirp
// (OFDM.py:259).  Each row copies [chirp | P x known | (L data symbols skipped) | P x known].
__global__ void __launch_bounds__(256) tx_frame_kernel(const FrameArgs a) {
    const int rows_per_stream = (int)a.pk_per_stream + 1;
    const int64_t row = blockIdx.x / a.gx;
    const int bx = blockIdx.x % a.gx;
    const int64_t stream = row / rows_per_stream;
    const int pk = (int)(row % rows_per_stream);
    const int64_t pkt_len = (int64_t)a.chirp_len + (int64_t)(2 * a.P + a.L) * a.symlen;
    float* o = a.out + stream * a.out_stride + (int64_t)pk * pkt_len;
    const int seg = (pk == (int)a.pk_per_stream) ? a.chirp_len : a.chirp_len + 2 * a.P * a.symlen;
    for (int i = bx * blockDim.x + threadIdx.x; i < seg; i += a.gx * blockDim.x) {
        if (i < a.chirp_len) {
            o[i] = a.chirp[i];
        } else {
            const int r = i - a.chirp_len;
            const int sidx = r / a.symlen, n = r - sidx * a.symlen;     // sidx in [0, 2P)
            const int slot = sidx < a.P ? sidx : sidx + a.L;
            o[a.chirp_len + (int64_t)slot * a.symlen + n] = a.known_time[n];
        }
    }
}

// sync_chirp (OFDM.py:106-109): cos(2 pi (f0 t + (f1-f0)/(2 t1) t^2)) * gain, t = linspace(0, t1, Lc)
__global__ void chirp_kernel(float* out, int Lc, double fs, double f0, double f1, double gain) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Lc) return;
    const double t1 = (double)Lc / fs;
    const double t = (double)n * (t1 / (double)(Lc - 1));
    const double beta = (f1 - f0) / t1;
    const double ph = f0 * t + 0.5 * beta * t * t;                    // in cycles
    out[n] = (float)(cospi(2.0 * (ph - floor(ph))) * gain);
}

int make_chirp(gf3_plan* plan) {
    const gf3_params& p = plan->p;
    GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp, (size_t)p.chirp_len * sizeof(float)));
    chirp_kernel<<<(p.chirp_len + 255) / 256, 256>>>(plan->d_chirp, p.chirp_len, (double)p.fs, (double)p.f0, (double)p.f1, (double)p.chirp_gain);
    GF3_LAUNCH_CHECK();
    GF3_CHECK_CUDA(cudaDeviceSynchronize());
    return GF3_OK;
}

template <class P>
static int launch_tx(const gf3_plan* plan, TxArgs a, const float* known, int64_t n_streams, float* known_time, cudaStream_t st) {
    constexpr int SF = kTxThreads / P::T;
    const gf3_params& p = plan->p;
    const int Nd = p.hi - p.lo;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (size_t)SF * ((2 * Nd + 7) / 8 + 1) + 16;
    // 1. the known symbol's time waveform (one symbol, gain applied) into scratch
    {
        TxArgs k = a;
        k.bits = nullptr; k.known = reinterpret_cast<const float2*>(known); k.out = known_time;
        auto kern = tx_symbols_kernel<P, true>;
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, kTxThreads, smem, st>>>(k);
        GF3_LAUNCH_CHECK();
    }
    // 2. data symbols
    {
        a.batches_per_packet = (a.L + SF - 1) / SF;
        const int64_t grid = n_streams * a.pk_per_stream * a.batches_per_packet;
        GF3_REQUIRE(grid <= 0x7fffffff, "tx_modulate: grid too large");
        auto kern = tx_symbols_kernel<P, false>;
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (grid > 0) {
            kern<<<(unsigned)grid, kTxThreads, smem, st>>>(a);
            GF3_LAUNCH_CHECK();
        }
    }
    // 3. chirps + known symbols
    {
        FrameArgs f;
        f.chirp = plan->d_chirp; f.known_time = known_time; f.out = a.out; f.out_stride = a.out_stride;
        f.pk_per_stream = a.pk_per_stream; f.n_streams = n_streams; f.symlen = p.N + p.cp; f.P = a.P; f.L = a.L;
        f.chirp_len = p.chirp_len;
        const int seg = p.chirp_len + 2 * a.P * f.symlen;
        int gx = (seg + 256 * 4 - 1) / (256 * 4);
        if (gx < 1) gx = 1;
        f.gx = gx;
        const int64_t rows = n_streams * (a.pk_per_stream + 1);
        GF3_REQUIRE(rows * gx <= 0x7fffffff, "tx_modulate: too many packets in one call");
        tx_frame_kernel<<<(unsigned)(rows * gx), 256, 0, st>>>(f);
        GF3_LAUNCH_CHECK();
    }
    return GF3_OK;
}

// Stage-level send_to_stream (gf3_stage.cu): every packet's [sync | P x known | .. | P x known] and the
// trailing sync, with the caller's own sync waveform (OFDM.py:244-259).
template <class P>
static int launch_frame_known(const gf3_plan* plan, const float* known, const float* sync, int sync_len,
                              int64_t n_packets, float* out, float* known_time, cudaStream_t st) {
    constexpr int SF = kTxThreads / P::T;
    const gf3_params& p = plan->p;
    const int Nd = p.hi - p.lo;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (size_t)SF * ((2 * Nd + 7) / 8 + 1) + 16;
    TxArgs k;
    memset(&k, 0, sizeof(k));
    k.known = reinterpret_cast<const float2*>(known); k.tw = plan->d_tw; k.out = known_time;
    k.cp = p.cp; k.lo = p.lo; k.hi = p.hi; k.P = p.n_pilots; k.L = p.packet_len; k.chirp_len = sync_len;
    k.gain = p.tx_gain / (float)p.N;
    auto kern = tx_symbols_kernel<P, true>;
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, kTxThreads, smem, st>>>(k);
    GF3_LAUNCH_CHECK();
    FrameArgs f;
    f.chirp = sync; f.known_time = known_time; f.out = out;
    f.pk_per_stream = n_packets; f.n_streams = 1; f.symlen = p.N + p.cp; f.P = p.n_pilots; f.L = p.packet_len;
    f.chirp_len = sync_len;
    f.out_stride = 0;
    const int seg = sync_len + 2 * f.P * f.symlen;
    int gx = (seg + 256 * 4 - 1) / (256 * 4);
    if (gx < 1) gx = 1;
    f.gx = gx;
    const int64_t rows = n_packets + 1;
    GF3_REQUIRE(rows * gx <= 0x7fffffff, "tx_frame: too many packets in one call");
    tx_frame_kernel<<<(unsigned)(rows * gx), 256, 0, st>>>(f);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

int tx_frame_known(const gf3_plan* plan, const float* known, const float* sync, int sync_len, int64_t n_packets,
                   float* out, cudaStream_t st) {
    const gf3_params& p = plan->p;
    float* known_time = const_cast<gf3_plan*>(plan)->d_known_time;
    if (!known_time) {
        GF3_CHECK_CUDA(cudaMalloc(&known_time, (size_t)(p.N + p.cp) * sizeof(float)));
        const_cast<gf3_plan*>(plan)->d_known_time = known_time;
    }
    switch (plan->logN) {
        case 6: return launch_frame_known<FftPlan<6>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 7: return launch_frame_known<FftPlan<7>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 8: return launch_frame_known<FftPlan<8>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 9: return launch_frame_known<FftPlan<9>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 10: return launch_frame_known<FftPlan<10>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 11: return launch_frame_known<FftPlan<11>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        case 12: return launch_frame_known<FftPlan<12>>(plan, known, sync, sync_len, n_packets, out, known_time, st);
        default: gf3::set_error("unsupported N"); return GF3_ERR_INVALID;
    }
}

}  // namespace gf3

using namespace gf3;

extern "C" int gf3_sync_chirp(const gf3_plan* plan, float* out, void* stream) {
    GF3_REQUIRE(plan && out, "sync_chirp: null argument");
    GF3_CHECK_CUDA(cudaMemcpyAsync(out, plan->d_chirp, (size_t)plan->p.chirp_len * sizeof(float),
                                   cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return GF3_OK;
}

extern "C" int gf3_tx_modulate(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride,
                               const float* filler, const float* known, int64_t n_streams,
                               int64_t pk_per_stream, float* out, int64_t out_stride, void* stream) {
    GF3_REQUIRE(plan && bits_packed && known && out, "tx_modulate: null argument");
    const gf3_params& p = plan->p;
    const int K = p.N / 2 - 1, Nd = p.hi - p.lo;
    GF3_REQUIRE(filler != nullptr || K == Nd, "tx_modulate: filler required when unused bins exist");
    GF3_REQUIRE(n_streams >= 0 && pk_per_stream >= 0, "tx_modulate: negative count");
    const int64_t need_bits = ((int64_t)p.packet_len * Nd * 2 + 7) / 8;
    GF3_REQUIRE(bits_stride >= need_bits, "tx_modulate: bits_stride %lld < %lld", (long long)bits_stride, (long long)need_bits);
    const int64_t pkt_len = (int64_t)p.chirp_len + (int64_t)(2 * p.n_pilots + p.packet_len) * (p.N + p.cp);
    GF3_REQUIRE(out_stride >= pkt_len * pk_per_stream + p.chirp_len, "tx_modulate: out_stride too small");
    if (n_streams == 0) return GF3_OK;
    TxArgs a;
    memset(&a, 0, sizeof(a));
    a.bits = bits_packed; a.filler = reinterpret_cast<const float2*>(filler);
    a.known = reinterpret_cast<const float2*>(known); a.tw = plan->d_tw; a.out = out;
    a.bits_stride = bits_stride; a.out_stride = out_stride; a.pk_per_stream = pk_per_stream > 0 ? pk_per_stream : 1;
    a.cp = p.cp; a.lo = p.lo; a.hi = p.hi; a.P = p.n_pilots; a.L = p.packet_len; a.chirp_len = p.chirp_len;
    a.gain = p.tx_gain / (float)p.N;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // scratch for the known symbol lives in the plan-independent per-call tail of `out`? No: keep
    // it inside the plan (allocated lazily, one symbol).
    float* known_time = const_cast<gf3_plan*>(plan)->d_known_time;
    if (!known_time) {
        GF3_CHECK_CUDA(cudaMalloc(&known_time, (size_t)(p.N + p.cp) * sizeof(float)));
        const_cast<gf3_plan*>(plan)->d_known_time = known_time;
    }
    if (pk_per_stream == 0) a.L = 0;
    a.pk_per_stream = pk_per_stream;
    switch (plan->logN) {
        case 6: return launch_tx<FftPlan<6>>(plan, a, known, n_streams, known_time, st);
        case 7: return launch_tx<FftPlan<7>>(plan, a, known, n_streams, known_time, st);
        case 8: return launch_tx<FftPlan<8>>(plan, a, known, n_streams, known_time, st);
        case 9: return launch_tx<FftPlan<9>>(plan, a, known, n_streams, known_time, st);
        case 10: return launch_tx<FftPlan<10>>(plan, a, known, n_streams, known_time, st);
        case 11: return launch_tx<FftPlan<11>>(plan, a, known, n_streams, known_time, st);
        case 12: return launch_tx<FftPlan<12>>(plan, a, known, n_streams, known_time, st);
        default: gf3::set_error("unsupported N"); return GF3_ERR_INVALID;
    }
}
