// gf3_rx.cu -- the receive chain of the GF3 modem as two fused sm_100a kernels.
//
//   rx_estimate_kernel : known-symbol channel estimate          (OFDM.py:407-418,593,429-462)
//   rx_demod_kernel    : CP strip + real FFT + one-tap equaliser + QPSK demap + XOR decode +
//                        bit packing, every data sample read once (OFDM.py:407-418,593,466-478,
//                        603,484-505,541-544); also the known-channel receiver of the
//                        Weekend-Challenge notebook (Weekend Challenge.ipynb:162-226)
//
// HBM-bound streaming work: no tensor cores, grids sized in waves of the SM count, coalesced
// 64/128-bit global access, FFT exchanges staged in shared memory.
#include "gf3_common.cuh"
#include "gf3_fft.cuh"

namespace gf3 {

constexpr int kThreads = 256;
// How the next FFT batch's samples are brought closer while the equaliser phase runs:
//   0 = nothing (plain loads at the start of the FFT phase)
//   1 = loads issued into registers before the equaliser phase
//   2 = one bulk L2 prefetch per symbol (cp.async.bulk.prefetch.L2, TMA engine, no registers)
#ifndef GF3_PREFETCH
#define GF3_PREFETCH 0
#endif
constexpr float kPi = 3.14159265358979323846f;

struct RxArgs {
    const float* samples;
    const int64_t* pkt_offset;   // may be null
    const float2* Hs;            // [n_packets, K]   (KNOWN_CH: Hinv[K], shared by all packets)
    const float2* He;            // [n_packets, K]
    const double* slope;         // [n_packets]
    const uint8_t* xor2;         // [Nd] or null
    uint8_t* bits;               // [n_packets, bits_stride] or null
    float2* eq;                  // [n_packets, L, K] or null
    const float2* tw;            // twiddle table (global)
    int64_t bits_stride;
    int64_t pkt_stride;          // (2P+L)(N+cp), used when pkt_offset == null
    int cp, lo, hi, P, L;
    int chunks_per_packet, chunks_per_cta, ctas_per_packet;
};

// streaming 8-byte load that does not pollute L1
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Load one symbol's N samples (CP already skipped by the caller) into the first-pass layout
// x[i] = z[t + i*T],  z[m] = (s[2m], s[2m+1]).
template <class P>
__device__ __forceinline__ void load_symbol(float2 (&x)[P::R], const float* __restrict__ s, int t) {
    if ((reinterpret_cast<uintptr_t>(s) & 7) == 0) {
#pragma unroll
        for (int i = 0; i < P::R; ++i) x[i] = ldg_stream2(s + 2 * (t + i * P::T));
    } else {   // odd sample offset (arbitrary sync index): two coalesced scalar loads
#pragma unroll
        for (int i = 0; i < P::R; ++i) {
            x[i].x = ldg_stream1(s + 2 * (t + i * P::T));
            x[i].y = ldg_stream1(s + 2 * (t + i * P::T) + 1);
        }
    }
}

// exp(-j * a) for a double-precision phase a (reduced in double, evaluated in float)
__device__ __forceinline__ float2 expmj(double a) {
    const double inv2pi = 0.15915494309189533577;
    double r = a * inv2pi;
    r -= rint(r);                       // revolutions in [-0.5, 0.5]
    float s, c;
    sincospif(2.0f * (float)r, &s, &c);
    return make_float2(c, -s);
}

// ------------------------------------------------------------------------------------------
// Data-symbol kernel.  One CTA owns a run of 16-symbol chunks of ONE packet.
//   phase A: SF symbols at a time, T threads per symbol, FFT in registers -> Z in smem
//   phase B: thread <-> bin pair (k, M-k): real-FFT untangling, equaliser, demap -> 2-bit codes
//   flush  : 16 codes -> one 32-bit word of MSB-first packed bits, coalesced store
// ------------------------------------------------------------------------------------------
template <class P, int NT, bool KNOWN_CH, bool WANT_EQ>
__global__ void __launch_bounds__(NT, 512 / NT) rx_demod_kernel(const RxArgs a) {
    constexpr int T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP;
    constexpr int SF = NT / T;                        // symbols per FFT batch
    constexpr int FLUSH = SF > 16 ? SF : 16;          // symbols per packed-bit flush (32*Nd bits: word aligned)
    constexpr int BATCHES = FLUSH / SF;
    constexpr int TB = (M / 2 < NT) ? M / 2 : NT;     // threads per symbol in phase B
    constexpr int SB = NT / TB;                       // symbols handled concurrently in phase B
    constexpr int PP = (M / 2) / TB;                  // bin pairs per thread
    constexpr int K = M - 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + SF * MP;
    uint8_t* stage = reinterpret_cast<uint8_t*>(tw + P::TW_TOTAL);            // [FLUSH*Nd] 2-bit codes, one per byte

    const int tid = threadIdx.x;
    const int64_t pkt = blockIdx.x / a.ctas_per_packet;
    const int c_first = (blockIdx.x % a.ctas_per_packet) * a.chunks_per_cta;
    const int c_last = min(c_first + a.chunks_per_cta, a.chunks_per_packet);
    const int Nd = a.hi - a.lo;
    const int L = a.L;

    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // XOR decode (OFDM.py:541-544) is applied per packed 32-bit word at flush time: word w of a
    // chunk covers codes 16w..16w+15, i.e. data carriers (16w+i) mod Nd -- the same for every chunk
    const bool use_xor = a.xor2 != nullptr;
    const int stage_bytes = ((FLUSH * Nd + 15) & ~15) + 16;
    uint32_t* xorw = reinterpret_cast<uint32_t*>(stage + stage_bytes);       // [FLUSH*Nd/16 + 1]
    if (use_xor && a.bits != nullptr) {
        const int wpc = (FLUSH * Nd + 15) >> 4;
        for (int w = tid; w < wpc; w += NT) {
            uint32_t word = 0;
            int c = (16 * w) % Nd;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (16 * w + i < FLUSH * Nd) {
                    const uint32_t code = a.xor2[c] & 3u;
                    word |= code << (8 * (i >> 2) + 6 - 2 * (i & 3));      // byte i/4, MSB-first inside the byte
                }
                c = (c + 1 == Nd) ? 0 : c + 1;
            }
            xorw[w] = word;
        }
    }

    const float* pkt_base = a.samples + (a.pkt_offset ? a.pkt_offset[pkt] : pkt * a.pkt_stride);
    const int symlen = N + a.cp;

    // ---- per-thread constants for phase B
    const int jb = tid % TB, sb = tid / TB;
    float2 w2[PP], u1[PP], u2[PP], G1[PP], G2[PP];
    int zo1[PP], zo2[PP], kk[PP];                       // padded smem offsets of Z[k], Z[M-k]; k itself
    const bool want_bits = a.bits != nullptr;
    int flags[PP];                                      // bit4 k is a data bin | bit5 km is a data bin | bit6 km exists (k != M/2)
    const float2* Hs = KNOWN_CH ? a.Hs : a.Hs + pkt * K;
    const double slope = KNOWN_CH ? 0.0 : a.slope[pkt];
    const double inv_lp = 1.0 / (double)(L + a.P);
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
        const int j = jb + pp * TB;
        const int k = j == 0 ? M / 2 : j, km = M - k;
        kk[pp] = k;
        zo1[pp] = zpad<P>(k);
        zo2[pp] = zpad<P>(km);
        float s, c;
        sincospif(2.0f * (float)k / (float)N, &s, &c);
        w2[pp] = make_float2(-s, -c);                   // -j * exp(-2 pi i k / N)
        int f = 0;
        if (k >= a.lo && k < a.hi) f |= 16;
        if (j != 0) {
            f |= 64;
            if (km >= a.lo && km < a.hi) f |= 32;
        }
        flags[pp] = f;
        if constexpr (!KNOWN_CH) {
            u1[pp] = expmj(slope * inv_lp * (double)((k - 1) * SB));
            u2[pp] = expmj(slope * inv_lp * (double)((km - 1) * SB));
        }
    }
    __syncthreads();

    // phase-A identity of this thread: symbol group ga, lane ta inside the group
    const int ga = tid / T, ta = tid % T;
    float2 x[R];
    auto load_batch = [&](int chunk, int b) {
        // symbols past the end of the packet are clamped to the last one: their spectra are
        // computed but never used (phase B only walks valid symbols), and no zero-fill is needed
        int l = chunk * FLUSH + b * SF + ga;
        l = l < L ? l : L - 1;
        load_symbol<P>(x, pkt_base + (int64_t)(a.P + l) * symlen + a.cp, ta);
    };
#if GF3_PREFETCH == 1
    load_batch(c_first, 0);
#endif

    for (int chunk = c_first; chunk < c_last; ++chunk) {
        const int l0 = chunk * FLUSH;
        const int nsym = min(FLUSH, L - l0);
        if ((nsym * Nd) & 15) {                       // last word of the chunk is partial: its missing codes are 0
            if (tid < 16) stage[nsym * Nd + tid] = 0;
        }
        // (re)seed the rotating equaliser taps exactly at the chunk start
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int j = jb + pp * TB;
            const int k = j == 0 ? M / 2 : j, km = M - k;
            const float2 h1 = Hs[k - 1], h2 = Hs[(j == 0 ? k : km) - 1];
            if constexpr (KNOWN_CH) {
                G1[pp] = h1;
                G2[pp] = h2;
            } else {
                const double wl = ((double)(l0 + sb) + 0.5 * (double)a.P) * inv_lp;   // OFDM.py:471,474
                G1[pp] = cmul(cconj(h1), expmj(slope * (double)(k - 1) * wl));
                G2[pp] = cmul(cconj(h2), expmj(slope * (double)(km - 1) * wl));
            }
        }

#pragma unroll 1
        for (int b = 0; b < BATCHES; ++b) {
            // ---------------- phase A: FFT of SF symbols
#if GF3_PREFETCH != 1
            load_batch(chunk, b);
#endif
            fft_forward<P, NT>(x, zbuf + ga * MP, tw, ta, ga);
            __syncthreads();
#if GF3_PREFETCH == 1
            // prefetch the next batch's samples: the loads fly while phase B computes
            load_batch(b + 1 < BATCHES ? chunk : chunk + 1, b + 1 < BATCHES ? b + 1 : 0);
#elif GF3_PREFETCH == 2
            if (ta == 0) {
                const int nc = b + 1 < BATCHES ? chunk : chunk + 1, nb = b + 1 < BATCHES ? b + 1 : 0;
                const int nl = nc * FLUSH + nb * SF + ga;
                if (nc < c_last && nl < L) {
                    const float* sp = pkt_base + (int64_t)(a.P + nl) * symlen + a.cp;
                    const uintptr_t lo16 = reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo16), "r"(N * 4 + 16) : "memory");
                }
            }
#endif
            // ---------------- phase B: untangle, equalise, demap
            // thread <-> PP bin pairs; walks the batch's symbols sb, sb+SB, ... with pointer increments
            {
                const int ls0 = b * SF + sb;                                   // first symbol (inside the chunk) of this thread
                int n_it = (nsym - ls0 + SB - 1) / SB;                          // valid symbols for this thread in this batch
                n_it = n_it < 0 ? 0 : (n_it > SF / SB ? SF / SB : n_it);
                const float2* zs = zbuf + sb * MP;
                uint8_t* st = stage + ls0 * Nd - a.lo;
                float2* eqp = nullptr;
                if constexpr (WANT_EQ) eqp = a.eq + ((int64_t)pkt * L + l0 + ls0) * K - 1;
#pragma unroll 2
                for (int it = 0; it < n_it; ++it) {
#pragma unroll
                    for (int pp = 0; pp < PP; ++pp) {
                        const float2 z1 = zs[zo1[pp]], z2 = zs[zo2[pp]];
                        const float2 s = make_float2(z1.x + z2.x, z1.y - z2.y);    // Z[k] + conj Z[M-k]
                        const float2 d = make_float2(z1.x - z2.x, z1.y + z2.y);    // Z[k] - conj Z[M-k]
                        const float2 tt = cmul(w2[pp], d);
                        const float2 x1 = cadd(s, tt);                             // 2 X[k]
                        const float2 x2 = make_float2(s.x - tt.x, tt.y - s.y);     // 2 X[M-k] = conj(s - tt)
                        const float2 y1 = cmul(x1, G1[pp]);
                        const float2 y2 = cmul(x2, G2[pp]);
                        if constexpr (!KNOWN_CH) {
                            G1[pp] = cmul(G1[pp], u1[pp]);
                            G2[pp] = cmul(G2[pp], u2[pp]);
                        }
                        const int f = flags[pp];
                        if (want_bits) {
                            if (f & 16) {
                                const unsigned code = ((__float_as_uint(y1.y) >> 30) & 2u) | (__float_as_uint(y1.x) >> 31);
                                st[kk[pp]] = (uint8_t)code;
                            }
                            if (f & 32) {
                                const unsigned code = ((__float_as_uint(y2.y) >> 30) & 2u) | (__float_as_uint(y2.x) >> 31);
                                st[M - kk[pp]] = (uint8_t)code;
                            }
                        }
                        if constexpr (WANT_EQ) {
                            const int k = kk[pp], km = M - k;
                            float sc1 = 0.5f, sc2 = 0.5f;
                            if constexpr (!KNOWN_CH) {
                                // |H| = |Hs| + (|He| - |Hs|) w  (OFDM.py:471); G carries conj(Hs) unnormalised
                                const float eq_w = (float)(((double)(l0 + ls0 + it * SB) + 0.5 * (double)a.P) * inv_lp);
                                const float2 hs1 = Hs[k - 1], he1 = a.He[pkt * K + k - 1];
                                const float a1 = sqrtf(hs1.x * hs1.x + hs1.y * hs1.y);
                                const float e1 = sqrtf(he1.x * he1.x + he1.y * he1.y);
                                sc1 = 0.5f / (a1 * (a1 + (e1 - a1) * eq_w));
                                if (f & 64) {
                                    const float2 hs2 = Hs[km - 1], he2 = a.He[pkt * K + km - 1];
                                    const float a2 = sqrtf(hs2.x * hs2.x + hs2.y * hs2.y);
                                    const float e2 = sqrtf(he2.x * he2.x + he2.y * he2.y);
                                    sc2 = 0.5f / (a2 * (a2 + (e2 - a2) * eq_w));
                                }
                            }
                            eqp[k] = make_float2(y1.x * sc1, y1.y * sc1);
                            if (f & 64) eqp[km] = make_float2(y2.x * sc2, y2.y * sc2);
                        }
                    }
                    zs += SB * MP;
                    st += SB * Nd;
                    if constexpr (WANT_EQ) eqp += (int64_t)SB * K;
                }
            }
            __syncthreads();
        }

        // ---------------- flush: 16 two-bit codes -> one 32-bit word (MSB-first bytes)
        // four codes c0..c3 (one per byte of u) -> (c0<<6 | c1<<4 | c2<<2 | c3) is the top byte of
        // u * 0x40100401 (no carries: every partial product lands on its own 2-bit field)
        if (want_bits) {
            const int ncodes = nsym * Nd;
            const int nwords = (ncodes + 15) >> 4;
            uint32_t* out = reinterpret_cast<uint32_t*>(a.bits + pkt * a.bits_stride) + (int64_t)l0 * Nd / 16;
            for (int w = tid; w < nwords; w += NT) {
                const uint4 v = *reinterpret_cast<const uint4*>(stage + 16 * w);
                const uint32_t t0 = v.x * 0x40100401u, t1 = v.y * 0x40100401u, t2 = v.z * 0x40100401u, t3 = v.w * 0x40100401u;
                uint32_t word = __byte_perm(__byte_perm(t0, t1, 0x0073), __byte_perm(t2, t3, 0x0073), 0x5410);
                if (use_xor) {
                    uint32_t xw = xorw[w];
                    const int r = ncodes - 16 * w;                  // codes in this word (< 16 only for the very last one)
                    if (r < 16) {                                   // keep the pad bits of a partial word zero
                        const int fb = r >> 2, rm = r & 3;
                        const uint32_t m = (fb ? (0xFFFFFFFFu >> (32 - 8 * fb)) : 0u) | (rm ? (((0xFF00u >> (2 * rm)) & 0xFFu) << (8 * fb)) : 0u);
                        xw &= m;
                    }
                    word ^= xw;
                }
                out[w] = word;
            }
            if (l0 + nsym >= L) {                                   // last chunk of the packet: clear the row's pad words
                const int stride_words = (int)(a.bits_stride / 4) - (int)((int64_t)l0 * Nd / 16);
                for (int w = nwords + tid; w < stride_words; w += NT) out[w] = 0u;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Channel-estimate kernel: one CTA per packet.
//   1. time-domain mean of the P leading / P trailing known symbols (FFT is linear, so
//      mean_P(FFT(x_p)) of OFDM.py:443-448 is computed as FFT(mean_P(x_p)): one FFT per block)
//   2. real FFT, divide by the known symbol                                     (OFDM.py:450-451)
//   3. phases -> unwrap along bins -> difference -> least-squares slope on the fit window
//      (OFDM.py:454-462), accumulated in double precision
// ------------------------------------------------------------------------------------------
struct EstArgs {
    const float* samples;
    const int64_t* pkt_offset;
    const float2* known;     // [K]
    float2* Hs;
    float2* He;
    double* slope;
    const float2* tw;
    int64_t pkt_stride;
    int cp, P, L, fit_lo, fit_hi;
};

template <class P>
__global__ void __launch_bounds__(kThreads) rx_estimate_kernel(const EstArgs a) {
    constexpr int NT = kThreads, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, K = M - 1;
    static_assert(2 * T <= NT, "estimate kernel needs both pilot blocks in one FFT batch");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* avg = reinterpret_cast<float2*>(smem_raw);            // [2][M]  (later: double phi[2][K])
    float2* zbuf = avg + 2 * M;                                    // [2][MP]
    float2* tw = zbuf + 2 * MP;
    double* phi = reinterpret_cast<double*>(smem_raw);            // aliases avg (dead after the FFT load)
    __shared__ int warp_tot[NT / 32];
    __shared__ double red[NT / 32];

    const int tid = threadIdx.x;
    const int64_t pkt = blockIdx.x;
    const float* pkt_base = a.samples + (a.pkt_offset ? a.pkt_offset[pkt] : pkt * a.pkt_stride);
    const int symlen = N + a.cp;

    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // ---- 1. time-domain sums of the pilot symbols (pure streaming: keep many 16-byte loads in flight)
    {
        const float* blk0 = pkt_base + a.cp;
        const float* blk1 = pkt_base + (int64_t)(a.P + a.L) * symlen + a.cp;
        const bool al16 = ((reinterpret_cast<uintptr_t>(blk0) | reinterpret_cast<uintptr_t>(blk1)) & 15) == 0 && (symlen % 4 == 0);
        if (al16) {
            constexpr int U = 10;
            for (int q = tid; q < 2 * (N / 4); q += NT) {          // float4 column q of block q / (N/4)
                const int blk = q / (N / 4), c4 = q % (N / 4);
                const float* s0 = (blk ? blk1 : blk0) + 4 * c4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p0 = 0; p0 < a.P; p0 += U) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int p = p0 + u < a.P ? p0 + u : a.P - 1;
                        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(s0 + (int64_t)p * symlen));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (p0 + u < a.P) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
                }
                *reinterpret_cast<float4*>(&avg[blk * M + 2 * c4]) = acc;
            }
        } else {
            for (int col = tid; col < 2 * M; col += NT) {
                const int blk = col / M, m = col % M;
                const float* s0 = (blk ? blk1 : blk0) + 2 * m;
                float2 acc = make_float2(0.f, 0.f);
                const bool al = (reinterpret_cast<uintptr_t>(s0) & 7) == 0 && (symlen % 2 == 0);
#pragma unroll 4
                for (int p = 0; p < a.P; ++p) {
                    const float* s = s0 + (int64_t)p * symlen;
                    float2 v;
                    if (al) v = ldg_stream2(s);
                    else { v.x = ldg_stream1(s); v.y = ldg_stream1(s + 1); }
                    acc.x += v.x;
                    acc.y += v.y;
                }
                avg[col] = acc;
            }
        }
    }
    __syncthreads();

    // ---- 2. FFT of both sums (warps holding an active symbol group take part as whole warps)
    constexpr int ACTIVE = (2 * T < 32) ? 32 : 2 * T;
    float2 x[R];
    if (tid < ACTIVE) {
        const int g = (tid / T) & 1, t = tid % T;
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = avg[g * M + t + i * T];
    }
    __syncthreads();                                   // avg is dead from here on (phi aliases it)
    if (tid < ACTIVE) {
        const int g = (tid / T) & 1, t = tid % T;
        fft_forward<P, ACTIVE>(x, zbuf + g * MP, tw, t, g);
    }
    __syncthreads();

    // ---- 3. untangle, divide by the known symbol, write Hs/He, phases to smem
    // Only phases inside the fit window are needed: np.unwrap's jumps before the window shift
    // unwrap(He) - unwrap(Hs) by a constant there, which does not change the fitted slope.
    const int flo = max(0, min(a.fit_lo, K)), fhi = max(flo, min(a.fit_hi, K));
    const float invP = 0.5f / (float)a.P;              // x1/x2 below are 2X
    for (int item = tid; item < 2 * (M / 2); item += NT) {
        const int blk = item / (M / 2), j = item % (M / 2);
        const int k = j == 0 ? M / 2 : j, km = M - k;
        const float2* zs = zbuf + blk * MP;
        float sn, cs;
        sincospif(2.0f * (float)k / (float)N, &sn, &cs);
        const float2 w2 = make_float2(-sn, -cs);
        const float2 z1 = zs[zpad<P>(k)], z2 = zs[zpad<P>(km)];
        const float2 s = make_float2(z1.x + z2.x, z1.y - z2.y);
        const float2 d = make_float2(z1.x - z2.x, z1.y + z2.y);
        const float2 tt = cmul(w2, d);
        const float2 x1 = cadd(s, tt);
        const float2 x2 = make_float2(s.x - tt.x, tt.y - s.y);
        float2* Hout = (blk ? a.He : a.Hs) + pkt * K;
        {
            const float2 kn = a.known[k - 1];           // |known| = 1: 1/known = conj(known)
            float2 h = cmul(x1, cconj(kn));
            h.x *= invP; h.y *= invP;
            Hout[k - 1] = h;
            if (k - 1 >= flo && k - 1 < fhi) phi[blk * K + k - 1] = atan2((double)h.y, (double)h.x);
        }
        if (j != 0) {
            const float2 kn = a.known[km - 1];
            float2 h = cmul(x2, cconj(kn));
            h.x *= invP; h.y *= invP;
            Hout[km - 1] = h;
            if (km - 1 >= flo && km - 1 < fhi) phi[blk * K + km - 1] = atan2((double)h.y, (double)h.x);
        }
    }
    __syncthreads();

    // ---- 4. unwrap both phase rows, difference, LS slope over [fit_lo, fit_hi) (0-based carrier index)
    const int nfit = fhi - flo;
    const int SEG = (nfit + NT - 1) / NT;
    const int i0 = flo + tid * SEG, i1 = min(fhi, i0 + SEG);
    const double PI = 3.14159265358979323846;
    // np.unwrap: a jump dd > pi subtracts 2 pi, dd < -pi adds 2 pi, |dd| == pi is left alone
    int local = 0;
    for (int i = max(i0, flo + 1); i < i1; ++i) {
        const double de = phi[K + i] - phi[K + i - 1], ds = phi[i] - phi[i - 1];
        local += (de > PI ? -1 : de < -PI ? 1 : 0) - (ds > PI ? -1 : ds < -PI ? 1 : 0);
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int prefix = incl - local;
    for (int w = 0; w < (tid >> 5); ++w) prefix += warp_tot[w];
    // walk the segment
    const double xbar = 0.5 * (double)(nfit - 1);
    double sxy = 0.0;
    int run = prefix;
    for (int i = i0; i < i1; ++i) {
        if (i >= flo + 1) {
            const double de = phi[K + i] - phi[K + i - 1], ds = phi[i] - phi[i - 1];
            run += (de > PI ? -1 : de < -PI ? 1 : 0) - (ds > PI ? -1 : ds < -PI ? 1 : 0);
        }
        if (i >= flo && i < fhi) {
            const double y = (phi[K + i] - phi[i]) + 2.0 * PI * (double)run;
            sxy += ((double)(i - flo) - xbar) * y;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    if ((tid & 31) == 0) red[tid >> 5] = sxy;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < NT / 32; ++w) tot += red[w];
        const double n = (double)nfit;
        const double sxx = n * (n * n - 1.0) / 12.0;
        a.slope[pkt] = nfit >= 2 ? tot / sxx : __longlong_as_double(0x7ff8000000000000LL);   // polyfit needs >= 2 points
    }
}

// ------------------------------------------------------------------------------------------ launchers
#ifndef GF3_DEMOD_THREADS
#define GF3_DEMOD_THREADS 128
#endif
// plans whose symbol group already spans >= 128 threads keep 256-thread CTAs (fewer bin pairs,
// hence less equaliser state, per thread)
template <class P> struct DemodThreads { static constexpr int value = P::T >= 128 ? 256 : GF3_DEMOD_THREADS; };

template <class P, bool KNOWN_CH, bool WANT_EQ>
static int launch_demod(const gf3_plan* plan, RxArgs a, int64_t n_packets, cudaStream_t st) {
    // CTA size: one symbol group needs P::T threads; smaller CTAs (more of them per SM) decorrelate
    // the load / FFT / equalise phases of co-resident CTAs
    constexpr int NT = (P::T > DemodThreads<P>::value) ? P::T : DemodThreads<P>::value;
    constexpr int SF = NT / P::T, FLUSH = SF > 16 ? SF : 16;
    const int Nd = a.hi - a.lo;
    a.chunks_per_packet = (a.L + FLUSH - 1) / FLUSH;
    // enough CTAs for ~8 waves of 2 CTAs/SM, otherwise one CTA walks the whole packet
    const int64_t want = (int64_t)plan->sm_count * 2 * 8;
    int64_t split = (want + n_packets - 1) / n_packets;
    if (split < 1) split = 1;
    if (split > a.chunks_per_packet) split = a.chunks_per_packet;
    a.chunks_per_cta = (int)((a.chunks_per_packet + split - 1) / split);
    a.ctas_per_packet = (a.chunks_per_packet + a.chunks_per_cta - 1) / a.chunks_per_cta;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (((size_t)FLUSH * Nd + 15) & ~(size_t)15) + 16
                        + (((size_t)FLUSH * Nd + 15) / 16 + 1) * sizeof(uint32_t);
    auto kern = rx_demod_kernel<P, NT, KNOWN_CH, WANT_EQ>;
    GF3_REQUIRE(smem <= 227 * 1024, "rx_demod: %zu bytes of shared memory needed (> 227 KB)", smem);
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = n_packets * a.ctas_per_packet;
    GF3_REQUIRE(grid <= 0x7fffffff, "rx_demod: grid too large");
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

template <class P>
static int launch_estimate(const gf3_plan* plan, EstArgs a, int64_t n_packets, cudaStream_t st) {
    const size_t smem = (size_t)(2 * P::M + 2 * P::MP + P::TW_TOTAL) * sizeof(float2);
    auto kern = rx_estimate_kernel<P>;
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)n_packets, kThreads, smem, st>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

#define GF3_DISPATCH_LOGN(logN, CALL)                                  \
    switch (logN) {                                                    \
        case 6: { using P = FftPlan<6>; CALL; } break;                 \
        case 7: { using P = FftPlan<7>; CALL; } break;                 \
        case 8: { using P = FftPlan<8>; CALL; } break;                 \
        case 9: { using P = FftPlan<9>; CALL; } break;                 \
        case 10: { using P = FftPlan<10>; CALL; } break;               \
        case 11: { using P = FftPlan<11>; CALL; } break;               \
        case 12: { using P = FftPlan<12>; CALL; } break;               \
        default: gf3::set_error("unsupported N = 2^%d", logN); return GF3_ERR_INVALID; \
    }

static int demod_common(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                        int64_t n_packets, const float* Hs, const float* He, const double* slope,
                        const uint8_t* xor2, uint8_t* bits, int64_t bits_stride, float* eq,
                        bool known_ch, int P_override, int L_override, void* stream) {
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
    if (known_ch) {
        if (eq) { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P, true, true>(plan, a, n_packets, st))); }
        else    { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P, true, false>(plan, a, n_packets, st))); }
    } else {
        if (eq) { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P, false, true>(plan, a, n_packets, st))); }
        else    { GF3_DISPATCH_LOGN(plan->logN, return (launch_demod<P, false, false>(plan, a, n_packets, st))); }
    }
    return GF3_OK;
}

}  // namespace gf3

using namespace gf3;

extern "C" int gf3_rx_estimate(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
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
    GF3_DISPATCH_LOGN(plan->logN, return (launch_estimate<P>(plan, a, n_packets, st)));
    return GF3_OK;
}

extern "C" int gf3_rx_demod(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                            int64_t n_packets, const float* Hs, const float* He, const double* slope,
                            const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                            void* stream) {
    return demod_common(plan, samples, pkt_offset, n_packets, Hs, He, slope, xor2, bits_packed,
                        bits_stride, eq, false, -1, -1, stream);
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
