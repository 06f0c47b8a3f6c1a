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
#define GF3_PREFETCH (-1)      // -1: per-plan default (see rx_demod_kernel)
#endif
#ifndef GF3_PK_L2_PREFETCH
#define GF3_PK_L2_PREFETCH 1
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
    int flush;                   // symbols per flush chunk
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

// 128-bit shared-memory load that the compiler may not split into (bank-conflicting) 32-bit loads
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
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
template <class P, int NT, int MINB, bool KNOWN_CH, bool WANT_EQ>
__global__ void __launch_bounds__(NT, MINB) rx_demod_kernel(const RxArgs a) {
    constexpr int T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP;
    constexpr int SF = NT / T;                        // symbols per FFT batch
    // symbols per packed-bit flush: a multiple of SF with FLUSH*Nd % 16 == 0, so every chunk starts
    // on a 32-bit word of the packet's bit stream (chosen by the launcher)
    const int FLUSH = a.flush;
    const int BATCHES = FLUSH / SF;
    constexpr int TB = (M / 2 < NT) ? M / 2 : NT;     // threads per symbol in phase B
    constexpr int SB = NT / TB;                       // symbols handled concurrently in phase B
    constexpr int PP = (M / 2) / TB;                  // bin pairs per thread
    constexpr int K = M - 1;
    // multi-warp symbol groups (N = 4096) have few loads per thread in flight and only two CTAs per
    // SM: issue the next batch's loads before the equaliser phase.  Half-warp groups (R = 32) have
    // no registers to spare for that and four CTAs per SM already overlap.
    constexpr int PREFETCH = (GF3_PREFETCH >= 0) ? GF3_PREFETCH : (T >= 128 ? 1 : 0);

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
    if constexpr (PREFETCH == 1) load_batch(c_first, 0);

    for (int chunk = c_first; chunk < c_last; ++chunk) {
        const int l0 = chunk * FLUSH;
        const int nsym = min(FLUSH, L - l0);
        if ((nsym * Nd) & 15) {                       // last word of the chunk is partial: its missing codes are 0
            if (tid < 16) stage[nsym * Nd + tid] = 0;
        }
        // (re)seed the rotating equaliser taps exactly (fp64 phase) at the first chunk and then every 64
        // symbols; in between the per-bin recurrence simply continues across chunk boundaries
        if (chunk == c_first || (l0 % 64) == 0)
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
            if constexpr (PREFETCH != 1) load_batch(chunk, b);
            fft_forward<P, NT>(x, zbuf + ga * MP, tw, ta, ga);
            __syncthreads();
            if constexpr (PREFETCH == 1) {
                // prefetch the next batch's samples: the loads fly while phase B computes
                load_batch(b + 1 < BATCHES ? chunk : chunk + 1, b + 1 < BATCHES ? b + 1 : 0);
            } else if constexpr (PREFETCH == 2) {
              if (ta == 0) {
                const int nc = b + 1 < BATCHES ? chunk : chunk + 1, nb = b + 1 < BATCHES ? b + 1 : 0;
                const int nl = nc * FLUSH + nb * SF + ga;
                if (nc < c_last && nl < L) {
                    const float* sp = pkt_base + (int64_t)(a.P + nl) * symlen + a.cp;
                    const uintptr_t lo16 = reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo16), "r"(N * 4 + 16) : "memory");
                }
              }
            }
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
// Packed data-symbol kernel (N = 1024 and N = 4096): same structure as rx_demod_kernel, but every
// symbol's M-point FFT is split by one scalar radix-2 DIF stage into two M/2-point FFTs that ride
// in the two lanes of FFMA2 / FADD2 / FMUL2, and the untangle / equalise / demap phase processes
// FOUR bins per thread per step: lanes (2j, 2j+1) and their mirrors (M-2j, M-2j-1).  That halves
// the FP32 issue slots of the kernel, which is instruction-issue bound (profiles/).
// ------------------------------------------------------------------------------------------
template <class P, int NT, int MINB, bool KNOWN_CH, bool WANT_EQ>
__global__ void __launch_bounds__(NT, MINB) rx_demod_pk_kernel(const RxArgs a) {
    // Warp-specialised: the first NT/2 threads (producers) load samples and run the FFTs of batch s
    // into Z buffer s&1; the other NT/2 threads (consumers) untangle / equalise / demap batch s-1 from
    // the other buffer and flush the packed bits.  They meet only on named barriers (full / empty per
    // buffer), so loads, FFT math and equaliser math of different batches overlap inside one CTA.
    constexpr int T = P::T, R = P::R, H = P::H, M = P::M, N = P::N, HP = P::HP, K = M - 1;
    constexpr int NP = NT / 2, NC = NT - NP;          // producer / consumer threads
    constexpr int SF = NP / T;                        // symbols per FFT batch
    constexpr int ITEMS = H / 2;                      // smem entries walked per symbol in phase B
    constexpr int TB = ITEMS < NC ? ITEMS : NC;       // consumer threads per symbol
    constexpr int SB = NC / TB;                       // symbols handled concurrently in phase B
    constexpr int PPK = ITEMS / TB;                   // entries per thread
    static_assert(SF >= 1 && SF % SB == 0, "bad producer/consumer split");
    enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4, BAR_CONS = 5 };
    constexpr int RESEED = 64;                        // symbols between exact re-seeds of the equaliser recurrence
    const int FLUSH = a.flush;
    const int BATCHES = FLUSH / SF;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* zbuf = reinterpret_cast<float4*>(smem_raw);                         // [2][SF][HP] (U.re, V.re, U.im, V.im)
    float2* tw = reinterpret_cast<float2*>(zbuf + 2 * SF * HP);                 // [TW_TOTAL]
    uint8_t* stage = reinterpret_cast<uint8_t*>(tw + P::TW_TOTAL);

    const int tid = threadIdx.x;
    const int64_t pkt = blockIdx.x / a.ctas_per_packet;
    const int c_first = (blockIdx.x % a.ctas_per_packet) * a.chunks_per_cta;
    const int c_last = min(c_first + a.chunks_per_cta, a.chunks_per_packet);
    const int Nd = a.hi - a.lo;
    const int L = a.L;

    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    const bool use_xor = a.xor2 != nullptr;
    const bool want_bits = a.bits != nullptr;
    const int stage_bytes = ((FLUSH * Nd + 15) & ~15) + 16;
    uint32_t* xorw = reinterpret_cast<uint32_t*>(stage + stage_bytes);
    if (use_xor && want_bits) {
        const int wpc = (FLUSH * Nd + 15) >> 4;
        for (int w = tid; w < wpc; w += NT) {
            uint32_t word = 0;
            int c = (16 * w) % Nd;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (16 * w + i < FLUSH * Nd) {
                    const uint32_t code = a.xor2[c] & 3u;
                    word |= code << (8 * (i >> 2) + 6 - 2 * (i & 3));
                }
                c = (c + 1 == Nd) ? 0 : c + 1;
            }
            xorw[w] = word;
        }
    }

    const float* pkt_base = a.samples + (a.pkt_offset ? a.pkt_offset[pkt] : pkt * a.pkt_stride);
    const int symlen = N + a.cp;

    // ---- per-thread constants for phase B.  Entry j carries bins (2j, 2j+1); their mirrors
    // (M-2j, M-2j-1) are lane x of entry H-j and lane y of entry H-1-j.
    const bool producer = tid < NP;
    const int ctid = producer ? 0 : tid - NP;          // consumer thread index
    const int jb = ctid % TB, sb = ctid / TB;
    cpk w2[PPK], u1[PPK], u2[PPK], G1[PPK], G2[PPK];
    int zo[PPK], zmx[PPK], zmy[PPK], jj[PPK], flags[PPK];   // flags: 1 own.x 2 own.y 4 mir.x 8 mir.y are data bins
    const float2* Hs = KNOWN_CH ? a.Hs : a.Hs + pkt * K;
    const double slope = KNOWN_CH ? 0.0 : a.slope[pkt];
    const double inv_lp = 1.0 / (double)(L + a.P);
    auto is_data = [&](int k) { return k >= a.lo && k < a.hi; };
    auto rot = [&](int k, double w) { return expmj(slope * (double)(k - 1) * w); };
#pragma unroll
    for (int pp = 0; pp < PPK; ++pp) {
        const int j = jb + pp * TB;
        jj[pp] = j;
        zo[pp] = zpad<P>(j);
        zmx[pp] = zpad<P>((H - j) % H);
        zmy[pp] = zpad<P>(H - 1 - j);
        const int kx = 2 * j, ky = 2 * j + 1, mx = M - 2 * j, my = M - 2 * j - 1;
        float sx, cx, sy, cy;
        sincospif(2.0f * (float)kx / (float)N, &sx, &cx);
        sincospif(2.0f * (float)ky / (float)N, &sy, &cy);
        w2[pp] = cpk{make_float2(-sx, -sy), make_float2(-cx, -cy)};          // -j exp(-2 pi i k / N)
        flags[pp] = (is_data(kx) ? 1 : 0) | (is_data(ky) ? 2 : 0) | ((j != 0 && is_data(mx)) ? 4 : 0) | (is_data(my) ? 8 : 0);
        if constexpr (!KNOWN_CH) {
            const double st = inv_lp * (double)SB;
            const float2 ax = rot(kx, st), ay = rot(ky, st), bx = rot(mx, st), by = rot(my, st);
            u1[pp] = cpk{make_float2(ax.x, ay.x), make_float2(ax.y, ay.y)};
            u2[pp] = cpk{make_float2(bx.x, by.x), make_float2(bx.y, by.y)};
        }
    }
    // bin M/2 (self-paired, = lane x of entry H/2) is walked by the thread that owns entry 0
    float2 Gh = make_float2(0.f, 0.f), uh = make_float2(1.f, 0.f);
    if constexpr (!KNOWN_CH) uh = rot(M / 2, inv_lp * (double)SB);
    const bool half_data = is_data(M / 2);
    __syncthreads();

    const int ga = tid / T, ta = tid % T;               // producer identity: symbol group, lane in the group
    const int n_chunks = c_last - c_first;
    const int n_steps = n_chunks * BATCHES;

    // 256-thread CTAs = two warpgroups: the producer warpgroup takes registers from the consumer one
    // (setmaxnreg), enough to keep the NEXT step's 2R loads in flight during the current FFT.
    constexpr bool DBUF = (NT == 256);
    if (producer) {
        // =============================== PRODUCERS: samples -> two packed H-point FFTs -> Z buffer
        if constexpr (DBUF) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        float2 za[R], zb[R];
        auto issue_loads = [&](int s) {
            const int chunk = c_first + s / BATCHES, b = s % BATCHES;
            int l = chunk * FLUSH + b * SF + ga;
            l = l < L ? l : L - 1;                          // clamped: spectra of invalid symbols are never used
            const float* sp = pkt_base + (int64_t)(a.P + l) * symlen + a.cp;
            if ((reinterpret_cast<uintptr_t>(sp) & 7) == 0) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    za[i] = ldg_stream2(sp + 2 * (ta + i * T));
                    zb[i] = ldg_stream2(sp + 2 * (ta + i * T + H));
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    za[i].x = ldg_stream1(sp + 2 * (ta + i * T));
                    za[i].y = ldg_stream1(sp + 2 * (ta + i * T) + 1);
                    zb[i].x = ldg_stream1(sp + 2 * (ta + i * T + H));
                    zb[i].y = ldg_stream1(sp + 2 * (ta + i * T + H) + 1);
                }
            }
        };
        if constexpr (DBUF) issue_loads(0);
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            if constexpr (!DBUF) {
                issue_loads(s);
                if (GF3_PK_L2_PREFETCH && ta == 0 && s + 1 < n_steps) {
                    // bulk L2 prefetch (TMA engine, no registers) of this group's symbol of the NEXT step
                    const int nc = c_first + (s + 1) / BATCHES, nb = (s + 1) % BATCHES;
                    const int nl = nc * FLUSH + nb * SF + ga;
                    if (nl < L) {
                        const float* np_ = pkt_base + (int64_t)(a.P + nl) * symlen + a.cp;
                        const uintptr_t lo16 = reinterpret_cast<uintptr_t>(np_) & ~(uintptr_t)15;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo16), "r"(N * 4 + 16) : "memory");
                    }
                }
            }
            // the loads are in flight while we wait for the consumers to release this buffer
            if (s >= 2) asm volatile("bar.sync %0, %1;" ::"r"((s & 1) ? BAR_EMPTY1 : BAR_EMPTY0), "n"(NT) : "memory");
            cpk x[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float2 u = cadd(za[i], zb[i]);
                const float2 v = cmul(csub(za[i], zb[i]), tw[ta + i * T]);      // W_M^h, h = ta + i*T
                x[i] = cpk{make_float2(u.x, v.x), make_float2(u.y, v.y)};
            }
            if constexpr (DBUF) {
                if (s + 1 < n_steps) issue_loads(s + 1);     // fly during the FFT below
            }
            pk_fft_forward<P, NP>(x, zbuf + ((s & 1) * SF + ga) * HP, tw, ta, ga);
            asm volatile("bar.arrive %0, %1;" ::"r"((s & 1) ? BAR_FULL1 : BAR_FULL0), "n"(NT) : "memory");
        }
        return;
    }
    if constexpr (DBUF) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");

    // =============================== CONSUMERS: untangle, equalise, demap, pack, store
#pragma unroll 1
    for (int chunk = c_first; chunk < c_last; ++chunk) {
        const int l0 = chunk * FLUSH;
        const int nsym = min(FLUSH, L - l0);
        if ((nsym * Nd) & 15) {
            if (ctid < 16) stage[nsym * Nd + ctid] = 0;
        }
        // (re)seed the rotating equaliser taps exactly (fp64 phase) at the first chunk and then every
        // RESEED symbols; in between the per-bin recurrence simply continues across chunk boundaries
        if (chunk == c_first || (l0 % RESEED) == 0) {
            const double wl = ((double)(l0 + sb) + 0.5 * (double)a.P) * inv_lp;          // OFDM.py:471,474
            auto seed = [&](int k) -> float2 {
                if (k < 1 || k > K) return make_float2(0.f, 0.f);
                const float2 h = Hs[k - 1];
                if constexpr (KNOWN_CH) return h;
                else return cmul(cconj(h), rot(k, wl));
            };
#pragma unroll
            for (int pp = 0; pp < PPK; ++pp) {
                const int j = jj[pp];
                const float2 gx = seed(2 * j), gy = seed(2 * j + 1), hx = seed(M - 2 * j), hy = seed(M - 2 * j - 1);
                G1[pp] = cpk{make_float2(gx.x, gy.x), make_float2(gx.y, gy.y)};
                G2[pp] = cpk{make_float2(hx.x, hy.x), make_float2(hx.y, hy.y)};
            }
            Gh = seed(M / 2);
        }

#pragma unroll 1
        for (int b = 0; b < BATCHES; ++b) {
            const int s = (chunk - c_first) * BATCHES + b;
            asm volatile("bar.sync %0, %1;" ::"r"((s & 1) ? BAR_FULL1 : BAR_FULL0), "n"(NT) : "memory");
            {
                const int ls0 = b * SF + sb;
                int n_it = (nsym - ls0 + SB - 1) / SB;
                n_it = n_it < 0 ? 0 : (n_it > SF / SB ? SF / SB : n_it);
                const float4* zs = zbuf + ((s & 1) * SF + sb) * HP;
                uint8_t* st = stage + ls0 * Nd - a.lo;
                float2* eqp = nullptr;
                if constexpr (WANT_EQ) eqp = a.eq + ((int64_t)pkt * L + l0 + ls0) * K - 1;
#pragma unroll 1
                for (int it = 0; it < n_it; ++it) {
#pragma unroll
                    for (int pp = 0; pp < PPK; ++pp) {
                        const float4 eo = lds128(zs + zo[pp]), ea = lds128(zs + zmx[pp]), eb = lds128(zs + zmy[pp]);
                        const cpk z1{make_float2(eo.x, eo.y), make_float2(eo.z, eo.w)};
                        const cpk z2{make_float2(ea.x, eb.y), make_float2(ea.z, eb.w)};
                        const cpk s{pk_add(z1.re, z2.re), pk_sub(z1.im, z2.im)};       // Z[k] + conj Z[M-k]
                        const cpk d{pk_sub(z1.re, z2.re), pk_add(z1.im, z2.im)};       // Z[k] - conj Z[M-k]
                        const cpk tt = cmul(w2[pp], d);
                        const cpk x1 = cadd(s, tt);                                     // 2 X[k]
                        const cpk x2{pk_sub(s.re, tt.re), pk_sub(tt.im, s.im)};         // 2 X[M-k] = conj(s - tt)
                        const cpk y1 = cmul(x1, G1[pp]);
                        const cpk y2 = cmul(x2, G2[pp]);
                        if constexpr (!KNOWN_CH) {
                            G1[pp] = cmul(G1[pp], u1[pp]);
                            G2[pp] = cmul(G2[pp], u2[pp]);
                        }
                        const int f = flags[pp];
                        const int kx = 2 * jj[pp];
                        if (want_bits) {
                            if (f & 1) st[kx] = (uint8_t)(((__float_as_uint(y1.im.x) >> 30) & 2u) | (__float_as_uint(y1.re.x) >> 31));
                            if (f & 2) st[kx + 1] = (uint8_t)(((__float_as_uint(y1.im.y) >> 30) & 2u) | (__float_as_uint(y1.re.y) >> 31));
                            if (f & 4) st[M - kx] = (uint8_t)(((__float_as_uint(y2.im.x) >> 30) & 2u) | (__float_as_uint(y2.re.x) >> 31));
                            if (f & 8) st[M - kx - 1] = (uint8_t)(((__float_as_uint(y2.im.y) >> 30) & 2u) | (__float_as_uint(y2.re.y) >> 31));
                        }
                        if constexpr (WANT_EQ) {
                            const float eq_w = KNOWN_CH ? 0.f : (float)(((double)(l0 + ls0 + it * SB) + 0.5 * (double)a.P) * inv_lp);
                            auto put = [&](int k, float yr, float yi) {
                                if (k < 1 || k > K) return;
                                float sc = 0.5f;
                                if constexpr (!KNOWN_CH) {
                                    const float2 hs = Hs[k - 1], he = a.He[pkt * K + k - 1];
                                    const float a1 = sqrtf(hs.x * hs.x + hs.y * hs.y), e1 = sqrtf(he.x * he.x + he.y * he.y);
                                    sc = 0.5f / (a1 * (a1 + (e1 - a1) * eq_w));       // |H| = |Hs| + (|He|-|Hs|) w  (OFDM.py:471)
                                }
                                eqp[k] = make_float2(yr * sc, yi * sc);
                            };
                            put(kx, y1.re.x, y1.im.x);
                            put(kx + 1, y1.re.y, y1.im.y);
                            if (kx != 0) put(M - kx, y2.re.x, y2.im.x);
                            put(M - kx - 1, y2.re.y, y2.im.y);
                        }
                    }
                    if (jb == 0) {      // bin M/2: 2 X[M/2] = 2 conj(Z[M/2]), Z[M/2] = U[H/2]
                        const float4 e = zs[zpad<P>(H / 2)];
                        const float2 xh = make_float2(2.f * e.x, -2.f * e.z);
                        const float2 yh = cmul(xh, Gh);
                        if constexpr (!KNOWN_CH) Gh = cmul(Gh, uh);
                        if (want_bits && half_data)
                            st[M / 2] = (uint8_t)(((__float_as_uint(yh.y) >> 30) & 2u) | (__float_as_uint(yh.x) >> 31));
                        if constexpr (WANT_EQ) {
                            float sc = 0.5f;
                            if constexpr (!KNOWN_CH) {
                                const float eq_w = (float)(((double)(l0 + ls0 + it * SB) + 0.5 * (double)a.P) * inv_lp);
                                const float2 hs = Hs[M / 2 - 1], he = a.He[pkt * K + M / 2 - 1];
                                const float a1 = sqrtf(hs.x * hs.x + hs.y * hs.y), e1 = sqrtf(he.x * he.x + he.y * he.y);
                                sc = 0.5f / (a1 * (a1 + (e1 - a1) * eq_w));
                            }
                            eqp[M / 2] = make_float2(yh.x * sc, yh.y * sc);
                        }
                    }
                    zs += SB * HP;
                    st += SB * Nd;
                    if constexpr (WANT_EQ) eqp += (int64_t)SB * K;
                }
            }
            if (s + 2 < n_steps) asm volatile("bar.arrive %0, %1;" ::"r"((s & 1) ? BAR_EMPTY1 : BAR_EMPTY0), "n"(NT) : "memory");
        }

        // ---------------- flush (consumer threads only)
        if (want_bits) {
            asm volatile("bar.sync %0, %1;" ::"r"((int)BAR_CONS), "n"(NC) : "memory");
            const int ncodes = nsym * Nd;
            const int nwords = (ncodes + 15) >> 4;
            uint32_t* out = reinterpret_cast<uint32_t*>(a.bits + pkt * a.bits_stride) + (int64_t)l0 * Nd / 16;
            for (int w = ctid; w < nwords; w += NC) {
                const uint4 v = *reinterpret_cast<const uint4*>(stage + 16 * w);
                const uint32_t t0 = v.x * 0x40100401u, t1 = v.y * 0x40100401u, t2 = v.z * 0x40100401u, t3 = v.w * 0x40100401u;
                uint32_t word = __byte_perm(__byte_perm(t0, t1, 0x0073), __byte_perm(t2, t3, 0x0073), 0x5410);
                if (use_xor) {
                    uint32_t xw = xorw[w];
                    const int r = ncodes - 16 * w;
                    if (r < 16) {
                        const int fb = r >> 2, rm = r & 3;
                        const uint32_t m = (fb ? (0xFFFFFFFFu >> (32 - 8 * fb)) : 0u) | (rm ? (((0xFF00u >> (2 * rm)) & 0xFFu) << (8 * fb)) : 0u);
                        xw &= m;
                    }
                    word ^= xw;
                }
                out[w] = word;
            }
            if (l0 + nsym >= L) {
                const int stride_words = (int)(a.bits_stride / 4) - (int)((int64_t)l0 * Nd / 16);
                for (int w = nwords + ctid; w < stride_words; w += NC) out[w] = 0u;
            }
            asm volatile("bar.sync %0, %1;" ::"r"((int)BAR_CONS), "n"(NC) : "memory");
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
// Plan used by the data-symbol kernel for each symbol size, its CTA size and CTAs per SM.
template <int LOGN> struct DemodCfg { using Plan = FftPlan<LOGN>; static constexpr int NT = GF3_DEMOD_THREADS, MINB = 512 / GF3_DEMOD_THREADS; };
// N = 4096: 128 threads per symbol (16 x 16 x 8), two symbols per 256-thread CTA.  (A warp-per-symbol
// 64 x 32 plan with ~255 registers / thread was measured slower: 8 warps per SM cannot hide latency.)
template <> struct DemodCfg<12> { using Plan = FftPlan<12>; static constexpr int NT = 256, MINB = 2; };

// Packed (two-lane) configuration per symbol size; Plan = void: use the scalar kernel.
template <int LOGN> struct DemodPkCfg { using Plan = void; static constexpr int NT = 128, MINB = 4; };
#ifndef GF3_NO_PACKED
#ifndef GF3_PK_NT
#define GF3_PK_NT 128
#define GF3_PK_MINB 4
#endif
template <> struct DemodPkCfg<10> { using Plan = PkPlan10; static constexpr int NT = GF3_PK_NT, MINB = GF3_PK_MINB; };
#endif

template <int LOGN, bool KNOWN_CH, bool WANT_EQ>
static int launch_demod_pk(const gf3_plan* plan, RxArgs a, int64_t n_packets, cudaStream_t st) {
    using P = typename DemodPkCfg<LOGN>::Plan;
    if constexpr (std::is_same<P, void>::value) {
        return GF3_ERR_INVALID;
    } else {
        constexpr int NT = (2 * P::T > DemodPkCfg<LOGN>::NT) ? 2 * P::T : DemodPkCfg<LOGN>::NT;
        constexpr int MINB = DemodPkCfg<LOGN>::MINB;
        constexpr int SF = (NT / 2) / P::T;             // producer half of the CTA
        const int Nd = a.hi - a.lo;
        int flush = SF;
        while ((flush * Nd) % 16 != 0 || flush < 8) flush += SF;
        a.flush = flush;
        a.tw = plan->d_tw_pk;
        a.chunks_per_packet = (a.L + flush - 1) / flush;
        const int64_t want = (int64_t)plan->sm_count * MINB * 8;
        int64_t split = (want + n_packets - 1) / n_packets;
        if (split < 1) split = 1;
        if (split > a.chunks_per_packet) split = a.chunks_per_packet;
        a.chunks_per_cta = (int)((a.chunks_per_packet + split - 1) / split);
        // CTA boundaries sit on multiples of the 64-symbol re-seed period, so the equaliser recurrence
        // (and hence every output bit) is independent of how a launch is split into CTAs
        { const int rc = flush >= 64 ? 1 : 64 / flush; a.chunks_per_cta = (a.chunks_per_cta + rc - 1) / rc * rc; }
        a.ctas_per_packet = (a.chunks_per_packet + a.chunks_per_cta - 1) / a.chunks_per_cta;
        const size_t smem = (size_t)2 * SF * P::HP * sizeof(float4) + (size_t)P::TW_TOTAL * sizeof(float2)
                            + (((size_t)flush * Nd + 15) & ~(size_t)15) + 16 + (((size_t)flush * Nd + 15) / 16 + 1) * sizeof(uint32_t);
        auto kern = rx_demod_pk_kernel<P, NT, MINB, KNOWN_CH, WANT_EQ>;
        GF3_REQUIRE(smem <= 227 * 1024, "rx_demod: %zu bytes of shared memory needed (> 227 KB)", smem);
        GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t grid = n_packets * a.ctas_per_packet;
        GF3_REQUIRE(grid <= 0x7fffffff, "rx_demod: grid too large");
        kern<<<(unsigned)grid, NT, smem, st>>>(a);
        GF3_LAUNCH_CHECK();
        return GF3_OK;
    }
}

template <int LOGN, bool KNOWN_CH, bool WANT_EQ>
static int launch_demod(const gf3_plan* plan, RxArgs a, int64_t n_packets, cudaStream_t st) {
    if constexpr (!std::is_same<typename DemodPkCfg<LOGN>::Plan, void>::value) {
        if (plan->d_tw_pk && plan->use_packed) return launch_demod_pk<LOGN, KNOWN_CH, WANT_EQ>(plan, a, n_packets, st);
    }
    using P = typename DemodCfg<LOGN>::Plan;
    // CTA size: one symbol group needs P::T threads; small CTAs (several per SM) decorrelate the
    // load / FFT / equalise phases of co-resident CTAs
    constexpr int NT = (P::T > DemodCfg<LOGN>::NT) ? P::T : DemodCfg<LOGN>::NT;
    constexpr int MINB = DemodCfg<LOGN>::MINB;
    constexpr int SF = NT / P::T;
    const int Nd = a.hi - a.lo;
    // smallest flush period: multiple of SF, FLUSH*Nd % 16 == 0, at least 8 symbols (amortise the flush)
    int flush = SF;
    while ((flush * Nd) % 16 != 0 || flush < 8) flush += SF;
    a.flush = flush;
    a.tw = plan->d_tw;
    a.chunks_per_packet = (a.L + flush - 1) / flush;
    // enough CTAs for ~8 waves, otherwise one CTA walks the whole packet
    const int64_t want = (int64_t)plan->sm_count * MINB * 8;
    int64_t split = (want + n_packets - 1) / n_packets;
    if (split < 1) split = 1;
    if (split > a.chunks_per_packet) split = a.chunks_per_packet;
    a.chunks_per_cta = (int)((a.chunks_per_packet + split - 1) / split);
    // CTA boundaries sit on multiples of the 64-symbol re-seed period, so the equaliser recurrence
    // (and hence every output bit) is independent of how a launch is split into CTAs
    { const int rc = flush >= 64 ? 1 : 64 / flush; a.chunks_per_cta = (a.chunks_per_cta + rc - 1) / rc * rc; }
    a.ctas_per_packet = (a.chunks_per_packet + a.chunks_per_cta - 1) / a.chunks_per_cta;
    const size_t smem = (size_t)(SF * P::MP + P::TW_TOTAL) * sizeof(float2) + (((size_t)flush * Nd + 15) & ~(size_t)15) + 16
                        + (((size_t)flush * Nd + 15) / 16 + 1) * sizeof(uint32_t);
    auto kern = rx_demod_kernel<P, NT, MINB, KNOWN_CH, WANT_EQ>;
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
