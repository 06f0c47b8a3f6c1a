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

#ifndef GF3_TX_THREADS
#define GF3_TX_THREADS 128
#endif
constexpr int kTxThreads = GF3_TX_THREADS;       // one symbol group needs P::T <= 128 threads; small CTAs decorrelate the phases
constexpr int kTxMinBlocks = 512 / GF3_TX_THREADS;

struct TxArgs {
    const uint8_t* bits;        // [n_streams, pk_per_stream, bits_stride]   (null for the known symbol)
    const float2* filler;       // [n_streams, K - Nd]
    const float2* known;        // [K] (used when bits == null); KNOWN_SYMBOL with n_work > 1: [n_work, K] spectra (gf3_tx_ifft)
    const uint8_t* xor2;        // [Nd] or null: (b0 << 1) | b1 of known_sequence[:2 Nd] -- encode("XOR") fused (OFDM.py:163-166)
    const float2* tw;
    float* out;                 // [n_streams, out_stride]
    int64_t bits_stride, out_stride;
    int64_t pk_per_stream;
    int cp, lo, hi, P, L, chirp_len;
    float gain;                 // tx_gain / N
    int batches_per_packet;     // ceil(L / SF)
    int64_t n_work;             // n_streams * pk_per_stream * batches_per_packet (persistent grid walks them)
};

// Persistent CTAs; work item = one batch of SF consecutive symbols of one packet.  Phase B' (thread <->
// bin pair (k, M-k), the same pairing as the receiver) builds conj(Z) in shared memory from the packed
// bits, phase A' runs the FFT, the epilogue writes x[2m] = Re Y[m], x[2m+1] = -Im Y[m] (times gain)
// plus the cyclic prefix.  Everything that depends only on the thread's bins (twiddle, bin class, bit
// position inside a symbol) is computed once per CTA.
template <class P, bool KNOWN_SYMBOL>
__global__ void __launch_bounds__(kTxThreads, kTxMinBlocks) tx_symbols_kernel(const TxArgs a) {
    constexpr int NT = kTxThreads, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, K = M - 1;
    constexpr int SF = NT / T;
    constexpr int TB = (M / 2 < NT) ? M / 2 : NT;     // threads per symbol in phase B'
    constexpr int SB = NT / TB;                       // symbols handled concurrently in phase B'
    constexpr int PP = (M / 2) / TB;                  // bin pairs per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + SF * MP;
    uint8_t* sbits = reinterpret_cast<uint8_t*>(tw + P::TW_TOTAL);     // [NB * NT] packed bits of the work item's symbols

    const int tid = threadIdx.x;
    const int Nd = a.hi - a.lo;
    const int symlen = N + a.cp;
    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // ---- per-thread constants of phase B': pairs k = jb + pp*TB (k = M/2 takes the slot of k = 0)
    const int jb = tid % TB, sb = tid / TB;
    float2 rot[PP];             // e^{+2 pi i k / N}
    int zo1[PP], zo2[PP];       // smem offsets of Z[k], Z[M-k]
    int src1[PP], src2[PP];     // bin k / M-k: >= 0 bit offset of the carrier's pair inside a symbol; -1 - u filler index; INT_MIN zero
    constexpr int kZero = (int)0x80000000;
    auto classify = [&](int kk) -> int {
        if (kk < 1 || kk > K) return kZero;                               // DC and Nyquist stay 0 (OFDM.py:209)
        if (KNOWN_SYMBOL) return -1 - (kk - 1);                           // "filler" = the known symbol itself
        if (kk >= a.lo && kk < a.hi) return 2 * (kk - a.lo);
        return -1 - (kk < a.lo ? kk - 1 : kk - 1 - Nd);                    // np.delete keeps ascending order (OFDM.py:49,213)
    };
    unsigned xk1[PP], xk2[PP];  // encode("XOR") of the thread's bins: the known bit pair moved onto bits 31 (b0) and 30 (b1)
    bool all_data = !KNOWN_SYMBOL;                    // every bin of this thread is a data carrier: no selects, no filler loads
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
        const int j = jb + pp * TB;
        const int k = j == 0 ? M / 2 : j, km = M - k;
        float sn, cs;
        sincospif(2.0f * (float)k / (float)N, &sn, &cs);
        rot[pp] = make_float2(cs, sn);
        src1[pp] = classify(k);
        src2[pp] = classify(km);
        all_data = all_data && src1[pp] >= 0 && src2[pp] >= 0;
        xk1[pp] = (a.xor2 && src1[pp] >= 0) ? ((unsigned)a.xor2[src1[pp] >> 1] & 3u) << 30 : 0u;
        xk2[pp] = (a.xor2 && src2[pp] >= 0) ? ((unsigned)a.xor2[src2[pp] >> 1] & 3u) << 30 : 0u;
        zo1[pp] = k;                                   // plain indexing: the hand-over to the FFT's register layout and the
        zo2[pp] = j != 0 ? km : 0;                     // natural-order output below are conflict-free without padding;            // the k = M/2 slot also clears Z[0] (DC and Nyquist: both 0)
    }
    const int g = tid / T, t = tid % T;

    const int64_t n_work = a.n_work;                  // KNOWN_SYMBOL: symbols given as spectra (1 for the known symbol itself)
    // The packed bits of a work item's symbols are one contiguous byte range of the packet (symbol l
    // starts at bit 2 Nd l).  They are fetched one item ahead into registers (NB bytes per thread), so
    // their latency is covered by the previous item's spectrum / FFT / store phases.
    constexpr int NB = R / 4 + 1;                     // >= ceil((SF * 2 Nd / 8 + 2) / NT), Nd < M
    uint8_t nb[NB];
    auto fetch_bits = [&](int64_t w) {
        const int64_t pg = w / a.batches_per_packet;
        const int lf = (int)(w % a.batches_per_packet) * SF;
        const int ns = min(SF, a.L - lf);
        const int64_t first = ((int64_t)lf * 2 * Nd) >> 3;
        int lim = (int)((((int64_t)(lf + ns) * 2 * Nd + 7) >> 3) - first);         // bytes of this item ...
        const int64_t room = a.bits_stride - first;                                // ... that exist in the row
        lim = room < lim ? (int)(room < 0 ? 0 : room) : lim;
        const uint8_t* pb = a.bits + pg * a.bits_stride + first + tid;
#pragma unroll
        for (int n = 0; n < NB; ++n) nb[n] = (tid + n * NT < lim) ? __ldg(pb + n * NT) : (uint8_t)0;
    };
    if constexpr (!KNOWN_SYMBOL) {
        if ((int64_t)blockIdx.x < n_work) fetch_bits(blockIdx.x);
    }
#pragma unroll 1
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
        int64_t pktg = 0;           // global packet index (stream * pk_per_stream + packet)
        int l_first = 0, nsym = 1;
        if constexpr (!KNOWN_SYMBOL) {
            pktg = work / a.batches_per_packet;
            l_first = (int)(work % a.batches_per_packet) * SF;
            nsym = min(SF, a.L - l_first);
        }
        const int64_t stream = pktg / a.pk_per_stream, pk = pktg % a.pk_per_stream;
        const float2* fill = KNOWN_SYMBOL ? a.known + work * K : a.filler + stream * (K - Nd);

        // stage the packed bits of the nsym symbols (bit offset l*2Nd is not byte aligned in general)
        if constexpr (!KNOWN_SYMBOL) {      // (the previous item's phase B' reads of sbits are two barriers back)
#pragma unroll
            for (int n = 0; n < NB; ++n) sbits[tid + n * NT] = nb[n];
        }
        __syncthreads();            // also: the previous work item's epilogue is done with zbuf
        if constexpr (!KNOWN_SYMBOL) {
            if (work + gridDim.x < n_work) fetch_bits(work + gridDim.x);
        }

        // ---- phase B': Hermitian spectrum -> conj(Z[k]) for the packed inverse real FFT
        //   E = X[k] + conj X[M-k],  O = (X[k] - conj X[M-k]) e^{+2 pi i k/N},  Z = E + jO
        for (int s = sb; s < nsym; s += SB) {              // (symbols past the packet's end: their buffers hold stale
            float2* zs = zbuf + s * MP;                    //  values, the FFT runs on them and nothing is written out)
            const int bit0 = s * 2 * Nd + (int)(((int64_t)l_first * 2 * Nd) & 7);   // first bit of symbol s in the staged range
            const uint8_t* sym = sbits;
            auto bin = [&](int src, unsigned xk) -> float2 {
                // data bin: QPSK of the encoded bit pair (OFDM.py:72-77), (b0,b1) -> ((1-2 b1) + j (1-2 b0)) / sqrt(2):
                // the two bits (MSB first) are moved onto the sign bits of +1/sqrt(2)
                const int bp = bit0 + (src > 0 ? src : 0);                 // even bit position inside the staged symbol
                const unsigned w = ((unsigned)sym[bp >> 3] << (24 + (bp & 7))) ^ xk;   // bit 31 = b0, bit 30 = b1
                float2 v = make_float2(__uint_as_float(0x3f3504f3u | ((w << 1) & 0x80000000u)),
                                       __uint_as_float(0x3f3504f3u | (w & 0x80000000u)));
                if (src < 0) v = (src == kZero) ? make_float2(0.f, 0.f) : fill[-1 - src];    // unused bin / filler (or known) symbol
                return v;
            };
            auto data_bin = [&](int bp, unsigned xk) -> float2 {         // bp: even bit position inside the staged range
                const unsigned w = ((unsigned)sym[bp >> 3] << (24 + (bp & 7))) ^ xk;   // bit 31 = b0, bit 30 = b1
                return make_float2(__uint_as_float(0x3f3504f3u | ((w << 1) & 0x80000000u)),
                                   __uint_as_float(0x3f3504f3u | (w & 0x80000000u)));
            };
            auto put_pair = [&](int pp, float2 X1, float2 X2) {
                const float2 E = make_float2(X1.x + X2.x, X1.y - X2.y);
                const float2 D = make_float2(X1.x - X2.x, X1.y + X2.y);
                const float2 O = cmul(D, rot[pp]);
                // Z[k] = E + jO ;  Z[M-k] = conj(E) + j conj(O); conj(Z) is stored (ifft = conj fft conj)
                const bool self = (jb + pp * TB) == 0;                     // k = M/2 pairs with itself
                zs[zo1[pp]] = make_float2(E.x - O.y, -(E.y + O.x));
                zs[zo2[pp]] = self ? make_float2(0.f, 0.f) : make_float2(E.x + O.y, E.y - O.x);
            };
            if (all_data) {
#pragma unroll
                for (int pp = 0; pp < PP; ++pp) put_pair(pp, data_bin(bit0 + src1[pp], xk1[pp]), data_bin(bit0 + src2[pp], xk2[pp]));
            } else {
#pragma unroll
                for (int pp = 0; pp < PP; ++pp) put_pair(pp, bin(src1[pp], xk1[pp]), bin(src2[pp], xk2[pp]));
            }
        }
        __syncthreads();

        // ---- phase A': forward FFT of conj(Z), last pass left in registers
        float2 x[R];
        {
            float2* zs = zbuf + g * MP;
#pragma unroll
            for (int i = 0; i < R; ++i) x[i] = zs[t + i * T];
            group_sync<P, NT>(g);
            fft_forward_to_regs<P, NT>(x, zs, tw, t, g);
        }

        // ---- epilogue: time samples with cyclic prefix (OFDM.py:221-226), gain (OFDM.py:256), straight from the
        // registers of the last pass: every symbol group writes its own symbol
        float* o0;
        if constexpr (KNOWN_SYMBOL) o0 = a.out + work * symlen;
        else o0 = a.out + stream * a.out_stride + pk * ((int64_t)a.chirp_len + (int64_t)(2 * a.P + a.L) * symlen)
                  + a.chirp_len + (int64_t)(a.P + l_first) * symlen;
        constexpr int LR = P::rad(P::NPASS - 1), LNS = P::ns(P::NPASS - 1), LQ = R / LR;       // last pass: radix, stride, sub-transforms
        constexpr bool PAIR = P::PAIRED || P::PAIRLAST;
        const int cp_first = (N - a.cp + 1) / 2;           // pairs m >= cp_first lie entirely inside the cyclic prefix's source
        if (g < nsym) {
            float* o = o0 + (int64_t)g * symlen;
            const uintptr_t amask = PAIR ? 15 : 7;
            const bool al = ((reinterpret_cast<uintptr_t>(o) | (uintptr_t)(a.cp * 4) | (uintptr_t)(symlen * 4)) & amask) == 0;
            if (al) {                                      // aligned symbol, even CP: coalesced vector stores
                if constexpr (PAIR) {
                    static_assert(!PAIR || LQ == 2, "paired last pass: two sub-transforms");
                    // x[i], x[LR + i] = points m = 2t + i*LNS and m + 1: four consecutive samples
                    float4* body = reinterpret_cast<float4*>(o + a.cp);
                    float4* pre = reinterpret_cast<float4*>(o - (N - a.cp));            // pre[m/2] = o[2m - (N - cp)]
#pragma unroll
                    for (int i = 0; i < LR; ++i) {
                        const int m = 2 * t + i * LNS;
                        const float4 v = make_float4(x[i].x * a.gain, -x[i].y * a.gain, x[LR + i].x * a.gain, -x[LR + i].y * a.gain);
                        body[m >> 1] = v;
                        if (m >= cp_first) pre[m >> 1] = v;          // (cp is a multiple of 4 here: m and m + 1 are on the same side)
                    }
                } else {
                    float2* body = reinterpret_cast<float2*>(o + a.cp);
                    float2* pre = reinterpret_cast<float2*>(o) - cp_first;      // pre[m] = o[2m - (N - cp)]
#pragma unroll
                    for (int q = 0; q < LQ; ++q) {
                        const int j = t + q * T;
                        const int base = (j / LNS) * (LNS * LR) + (j % LNS);
#pragma unroll
                        for (int i = 0; i < LR; ++i) {
                            const int m = base + i * LNS;
                            const float2 v = make_float2(x[q * LR + i].x * a.gain, -x[q * LR + i].y * a.gain);
                            body[m] = v;
                            if (m >= cp_first) pre[m] = v;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < LQ; ++q) {
                    const int j = PAIR ? 2 * t + q : t + q * T;
                    const int base = (j / LNS) * (LNS * LR) + (j % LNS);
#pragma unroll
                    for (int i = 0; i < LR; ++i) {
                        const int m = base + i * LNS;
                        const float v0 = x[q * LR + i].x * a.gain, v1 = -x[q * LR + i].y * a.gain;
                        const int n0 = 2 * m - (N - a.cp);                   // position of this pair inside the cyclic prefix
                        o[a.cp + 2 * m] = v0;
                        o[a.cp + 2 * m + 1] = v1;
                        if (n0 >= 0) o[n0] = v0;
                        if (n0 + 1 >= 0) o[n0 + 1] = v1;
                    }
                }
            }
        }
        // (the next work item's first barrier orders this item's FFT exchanges before its phase B' writes)
    }
}

// Chirp preambles and known symbols: plain copies.
struct FrameArgs {
    const float* chirp;         // [chirp_len]
    const float* known_time;    // [N + cp], gain applied
    float* out;
    int64_t out_stride, pk_per_stream, n_streams;
    int symlen, P, L, chirp_len;
};

// One CTA copies one segment of one row: row = stream * (pk_per_stream + 1) + packet (the extra "packet"
// is the trailing chirp, OFDM.py:259); segment 0 = the chirp, segments 1 .. 2P = the known symbols
// before and after the L data symbols (which tx_symbols_kernel writes).  No per-sample index math.
__global__ void __launch_bounds__(256) tx_frame_kernel(const FrameArgs a) {
    const unsigned row = blockIdx.x;                                     // grid = (rows, a few CTAs sharing the 2P + 1 segments)
    const unsigned rows_per_stream = (unsigned)a.pk_per_stream + 1u;
    const int64_t stream = row / rows_per_stream;
    const int pk = (int)(row % rows_per_stream);
    const int64_t pkt_len = (int64_t)a.chirp_len + (int64_t)(2 * a.P + a.L) * a.symlen;
    float* const row_out = a.out + stream * a.out_stride + (int64_t)pk * pkt_len;
    const int nseg = pk == (int)a.pk_per_stream ? 1 : 2 * a.P + 1;      // the trailing row is a chirp only
    for (int seg = blockIdx.y; seg < nseg; seg += gridDim.y) {
    float* o = row_out;
    const float* src;
    int n;
    if (seg == 0) { src = a.chirp; n = a.chirp_len; }
    else {
        const int sidx = seg - 1;                                       // 0 .. 2P-1
        const int slot = sidx < a.P ? sidx : sidx + a.L;
        o += a.chirp_len + (int64_t)slot * a.symlen;
        src = a.known_time;
        n = a.symlen;
        if (o == src) continue;                                         // the slot the waveform was computed into
    }
    if (((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(o)[i] = reinterpret_cast<const float4*>(src)[i];
        for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) o[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = src[i];
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

// CTAs per frame row: enough CTAs to fill the GPU a few times over, each looping over its share of
// the row's 2P + 1 segments (one CTA per 4 KB segment was mostly launch overhead)
static unsigned frame_grid_y(const gf3_plan* plan, int64_t rows, int64_t nseg) {
    int64_t gy = ((int64_t)plan->sm_count * 16 + rows - 1) / rows;
    gy = gy < 1 ? 1 : (gy > nseg ? nseg : gy);
    return (unsigned)(gy > 65535 ? 65535 : gy);
}

// The known symbol's time waveform is computed once per call straight into the FIRST known-symbol slot
// of the output (stream 0, packet 0) and copied from there into every other slot by tx_frame_kernel:
// no scratch in the (immutable, shareable) plan, so concurrent calls on different streams do not meet.
template <class P>
static int launch_tx(const gf3_plan* plan, TxArgs a, const float* known, int64_t n_streams, cudaStream_t st) {
    float* const known_time = a.out + a.chirp_len;
    constexpr int SF = kTxThreads / P::T;
    const gf3_params& p = plan->p;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (size_t)(P::R / 4 + 1) * kTxThreads + 16;   // staged bits: NB bytes per thread
    // 1. the known symbol's time waveform (one symbol, gain applied)
    if (a.P > 0 && a.pk_per_stream > 0) {
        TxArgs k = a;
        k.bits = nullptr; k.xor2 = nullptr; k.n_work = 1; k.known = reinterpret_cast<const float2*>(known); k.out = known_time;
        auto kern = tx_symbols_kernel<P, true>;
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, kTxThreads, smem, st>>>(k);
        GF3_LAUNCH_CHECK();
    }
    // 2. data symbols
    {
        a.batches_per_packet = (a.L + SF - 1) / SF;
        a.n_work = n_streams * a.pk_per_stream * a.batches_per_packet;
        auto kern = tx_symbols_kernel<P, false>;
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GF3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTxThreads, smem));
        int64_t grid = (int64_t)plan->sm_count * (per_sm < 1 ? 1 : per_sm);         // persistent: one full wave
        if (grid > a.n_work) grid = a.n_work;
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
        const int64_t rows = n_streams * (a.pk_per_stream + 1);
        const int64_t nseg = 2 * a.P + 1;
        GF3_REQUIRE(rows <= 0x7fffffff, "tx_modulate: too many packets in one call");
        tx_frame_kernel<<<dim3((unsigned)rows, frame_grid_y(plan, rows, nseg)), 256, 0, st>>>(f);
        GF3_LAUNCH_CHECK();
    }
    return GF3_OK;
}

// Stage-level send_to_stream (gf3_stage.cu): every packet's [sync | P x known | .. | P x known] and the
// trailing sync, with the caller's own sync waveform (OFDM.py:244-259).
template <class P>
static int launch_frame_known(const gf3_plan* plan, const float* known, const float* sync, int sync_len,
                              int64_t n_packets, float* out, cudaStream_t st) {
    float* const known_time = out + sync_len;      // first known-symbol slot of the frame (see launch_tx)
    constexpr int SF = kTxThreads / P::T;
    const gf3_params& p = plan->p;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (size_t)(P::R / 4 + 1) * kTxThreads + 16;   // staged bits: NB bytes per thread
    TxArgs k;
    memset(&k, 0, sizeof(k));
    k.known = reinterpret_cast<const float2*>(known); k.tw = plan->d_tw; k.out = known_time;
    k.cp = p.cp; k.lo = p.lo; k.hi = p.hi; k.P = p.n_pilots; k.L = p.packet_len; k.chirp_len = sync_len;
    k.gain = p.tx_gain / (float)p.N;
    k.n_work = 1; k.pk_per_stream = 1;
    if (p.n_pilots > 0 && n_packets > 0) {
        auto kern = tx_symbols_kernel<P, true>;
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, kTxThreads, smem, st>>>(k);
        GF3_LAUNCH_CHECK();
    }
    FrameArgs f;
    f.chirp = sync; f.known_time = known_time; f.out = out;
    f.pk_per_stream = n_packets; f.n_streams = 1; f.symlen = p.N + p.cp; f.P = p.n_pilots; f.L = p.packet_len;
    f.chirp_len = sync_len;
    f.out_stride = 0;
    const int64_t rows = n_packets + 1;
    const int64_t nseg = 2 * f.P + 1;
    GF3_REQUIRE(rows <= 0x7fffffff, "tx_frame: too many packets in one call");
    tx_frame_kernel<<<dim3((unsigned)rows, frame_grid_y(plan, rows, nseg)), 256, 0, st>>>(f);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

int tx_frame_known(const gf3_plan* plan, const float* known, const float* sync, int sync_len, int64_t n_packets,
                   float* out, cudaStream_t st) {
    switch (plan->logN) {
        case 6: return launch_frame_known<FftPlan<6>>(plan, known, sync, sync_len, n_packets, out, st);
        case 7: return launch_frame_known<FftPlan<7>>(plan, known, sync, sync_len, n_packets, out, st);
        case 8: return launch_frame_known<FftPlan<8>>(plan, known, sync, sync_len, n_packets, out, st);
        case 9: return launch_frame_known<FftPlan<9>>(plan, known, sync, sync_len, n_packets, out, st);
        case 10: return launch_frame_known<FftPlan<10>>(plan, known, sync, sync_len, n_packets, out, st);
        case 11: return launch_frame_known<FftPlan<11>>(plan, known, sync, sync_len, n_packets, out, st);
        case 12: return launch_frame_known<FftPlan<12>>(plan, known, sync, sync_len, n_packets, out, st);
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

static int tx_modulate_common(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride, const uint8_t* xor2,
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
    a.bits = bits_packed; a.xor2 = xor2; a.filler = reinterpret_cast<const float2*>(filler);
    a.known = reinterpret_cast<const float2*>(known); a.tw = plan->d_tw; a.out = out;
    a.bits_stride = bits_stride; a.out_stride = out_stride; a.pk_per_stream = pk_per_stream > 0 ? pk_per_stream : 1;
    a.cp = p.cp; a.lo = p.lo; a.hi = p.hi; a.P = p.n_pilots; a.L = p.packet_len; a.chirp_len = p.chirp_len;
    a.gain = p.tx_gain / (float)p.N;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (pk_per_stream == 0) a.L = 0;
    a.pk_per_stream = pk_per_stream;
    switch (plan->logN) {
        case 6: return launch_tx<FftPlan<6>>(plan, a, known, n_streams, st);
        case 7: return launch_tx<FftPlan<7>>(plan, a, known, n_streams, st);
        case 8: return launch_tx<FftPlan<8>>(plan, a, known, n_streams, st);
        case 9: return launch_tx<FftPlan<9>>(plan, a, known, n_streams, st);
        case 10: return launch_tx<FftPlan<10>>(plan, a, known, n_streams, st);
        case 11: return launch_tx<FftPlan<11>>(plan, a, known, n_streams, st);
        case 12: return launch_tx<FftPlan<12>>(plan, a, known, n_streams, st);
        default: gf3::set_error("unsupported N"); return GF3_ERR_INVALID;
    }
}

extern "C" int gf3_tx_modulate(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride,
                               const float* filler, const float* known, int64_t n_streams,
                               int64_t pk_per_stream, float* out, int64_t out_stride, void* stream) {
    return tx_modulate_common(plan, bits_packed, bits_stride, nullptr, filler, known, n_streams, pk_per_stream, out, out_stride, stream);
}

extern "C" int gf3_tx_encode_modulate(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride, const uint8_t* xor2,
                                      const float* filler, const float* known, int64_t n_streams,
                                      int64_t pk_per_stream, float* out, int64_t out_stride, void* stream) {
    GF3_REQUIRE(xor2 != nullptr, "tx_encode_modulate: null xor2");
    return tx_modulate_common(plan, bits_packed, bits_stride, xor2, filler, known, n_streams, pk_per_stream, out, out_stride, stream);
}

template <class P>
static int launch_ifft(const gf3_plan* plan, const float2* spec, int64_t n, float* out, cudaStream_t st) {
    constexpr int SF = kTxThreads / P::T;
    const gf3_params& p = plan->p;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (size_t)(P::R / 4 + 1) * kTxThreads + 16;
    TxArgs k;
    memset(&k, 0, sizeof(k));
    k.known = spec; k.tw = plan->d_tw; k.out = out;
    k.cp = p.cp; k.lo = p.lo; k.hi = p.hi; k.P = 0; k.L = 1; k.chirp_len = 0;
    k.gain = 1.0f / (float)p.N;                        // np.fft.ifft scaling, no transmit gain
    k.n_work = n; k.pk_per_stream = 1;
    auto kern = tx_symbols_kernel<P, true>;
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (int64_t)plan->sm_count * kTxMinBlocks;
    if (grid > n) grid = n;
    kern<<<(unsigned)grid, kTxThreads, smem, st>>>(k);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_tx_ifft(const gf3_plan* plan, const float* spectrum, int64_t n_symbols, float* out, void* stream) {
    GF3_REQUIRE(plan && spectrum && out, "tx_ifft: null argument");
    GF3_REQUIRE(n_symbols >= 0, "tx_ifft: negative count");
    if (n_symbols == 0) return GF3_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float2* spec = reinterpret_cast<const float2*>(spectrum);
    switch (plan->logN) {
        case 6: return launch_ifft<FftPlan<6>>(plan, spec, n_symbols, out, st);
        case 7: return launch_ifft<FftPlan<7>>(plan, spec, n_symbols, out, st);
        case 8: return launch_ifft<FftPlan<8>>(plan, spec, n_symbols, out, st);
        case 9: return launch_ifft<FftPlan<9>>(plan, spec, n_symbols, out, st);
        case 10: return launch_ifft<FftPlan<10>>(plan, spec, n_symbols, out, st);
        case 11: return launch_ifft<FftPlan<11>>(plan, spec, n_symbols, out, st);
        case 12: return launch_ifft<FftPlan<12>>(plan, spec, n_symbols, out, st);
        default: gf3::set_error("unsupported N"); return GF3_ERR_INVALID;
    }
}
