// gf3_rx_kernels.cuh -- the receive chain of the GF3 modem as fused sm_100a kernels.
//
//   rx_estimate_kernel : known-symbol channel estimate          (OFDM.py:407-418,593,429-462)
//   rx_demod_kernel    : CP strip + real FFT + one-tap equaliser + QPSK demap + XOR decode +
//                        bit packing, every data sample read once (OFDM.py:407-418,593,466-478,
//                        603,484-505,541-544); also the known-channel receiver of the
//                        Weekend-Challenge notebook (Weekend Challenge.ipynb:162-226)
//   rx_demod_kernel<.., FUSE_EST> : both in ONE launch (gf3_rx_receive): a persistent CTA estimates a
//                        packet's channel when it first touches the packet (estimate_packet)
//
// Streaming work with no dense contraction: no tensor cores, persistent grids of one full wave,
// coalesced 64/128-bit global access, FFT exchanges staged in shared memory, f32x2 (FADD2 / FMUL2 /
// FFMA2) arithmetic in the butterflies and in the bin-pair phase.
//
// Compile-time knobs (experiments; the defaults are what profiles/ measured as fastest):
//   GF3_PREFETCH       how the next FFT batch's samples are brought closer (see below)
//   GF3_EST_U          pilot loads in flight per thread in the fused estimate
//   GF3_PHASEB_UNROLL  symbols per unrolled step of the bin-pair phase
//   GF3_FLUSH_UNROLL   packed words per unrolled step of the flush
//   GF3_DEMOD_NATURAL  1: the last FFT pass leaves the spectrum unpadded (no mirrored-read conflicts)
//   GF3_FUSE_SEQUENTIAL 1: fuse the estimate at N = 4096 too (measured slower: 0.80 vs 0.71 ms on C4)
//   GF3_DMASK_ONCE     data-carrier mask of a thread's bins: once per CTA (1) or per batch (0); -1 = per plan
//   GF3_DEMOD_THREADS / GF3_DEMOD11_THREADS / GF3_DEMOD12_THREADS (+ _MINB)  CTA size and CTAs per SM
//                      for N <= 1024 (128 x 4), N = 2048 (256 x 2), N = 4096 (256 x 2)
//   GF3_ABL            ablation mask for timing only (1: no bin-pair phase, 2: no FFT, 4: no global
//                      loads, 8: no code stores); results are wrong with any bit set
#pragma once
#include <stdlib.h>

#include "gf3_common.cuh"
#include "gf3_fft.cuh"
#include "gf3_fit.cuh"

namespace gf3 {

constexpr int kThreads = 256;
// How the next FFT batch's samples are brought closer while the equaliser phase runs:
//   0 = nothing (plain loads at the start of the FFT phase)
//   1 = loads issued into registers before the equaliser phase
//   2 = one bulk L2 prefetch per symbol (cp.async.bulk.prefetch.L2, TMA engine, no registers)
#ifndef GF3_PREFETCH
#define GF3_PREFETCH (-1)      // -1: per-plan default (see rx_demod_kernel)
#endif
#ifndef GF3_PREFETCH_PART
#define GF3_PREFETCH_PART 16   // GF3_PREFETCH == 3: registers (complex points) of the next batch requested ahead
#endif
#ifndef GF3_WHOLE_PACKETS
#define GF3_WHOLE_PACKETS 0    // 1: CTA ranges on packet boundaries in the fused kernel (no packet estimated twice): C3 0.976 vs 0.981 ms in
#endif                         // one run, 1.035 in the next -- every CTA then sits in the same phase of its packet at the same time

#ifndef GF3_EST_U
#define GF3_EST_U 20
#endif
#ifndef GF3_FUSE_SEQUENTIAL
#define GF3_FUSE_SEQUENTIAL 0      // 1: fuse the estimate at N = 4096 too (pilot blocks one after the other)
#endif
#ifndef GF3_FLUSH_UNROLL
#define GF3_FLUSH_UNROLL 0     // 0: per plan (4 words per step at N <= 1024: C3 0.978 vs 0.981 ms; 2 above: A2 0.684 vs 0.687)
#endif
#ifndef GF3_DMASK_ONCE
#define GF3_DMASK_ONCE (-1)    // data-carrier mask of a thread's bins: 1 once per CTA, 0 per batch, -1 per-plan default
#endif
#ifndef GF3_PHASEB_UNROLL
#define GF3_PHASEB_UNROLL 8
#endif
#ifndef GF3_DEMOD_NATURAL
#define GF3_DEMOD_NATURAL 1
#endif
#ifndef GF3_ABL
#define GF3_ABL 0
#endif

struct RxArgs {
    const void* samples;         // float32 / int16 / uint8 samples (template parameter S of the kernels)
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
    int64_t n_packets;
    int chunks_per_packet;       // flush chunks (work items) per packet = ceil(L / flush)
    int flush;                   // symbols per flush chunk
    // fused channel estimate (FUSE_EST): Hs / He / slope above are then OUTPUTS of the same launch
    const float2* known;         // [K]
    int fit_lo, fit_hi;
};

// streaming 8-byte load that does not pollute L1
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
// Received samples in their recorded format.  uint8 PCM carries a DC offset of 128 that the reference keeps
// (Final System Test.ipynb:86, `r/1.0`); a constant only lands in FFT bin 0, which the receive chain never
// uses, so it is removed here (exactly: small integers) -- the float32 FFT then rounds relative to the
// signal's energy instead of the offset's.
template <class S> __device__ __forceinline__ float cvt_sample(S v) { return (float)v; }
template <> __device__ __forceinline__ float cvt_sample<uint8_t>(uint8_t v) { return (float)((int)v - 128); }
// Narrow loads of the element-aligned-only paths are written in PTX: left to the compiler, the two single-sample
// loads of a pair get merged into one wider load that faults at odd sample offsets.
__device__ __forceinline__ int ldg_elem(const int16_t* p) { int v; asm volatile("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ldg_elem(const uint8_t* p) { int v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ldg_elem(const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ int lds_s16(const unsigned char* p) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p))); return v; }
__device__ __forceinline__ int lds_u8(const unsigned char* p) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p))); return v; }
__device__ __forceinline__ float lds_f32(const unsigned char* p) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p))); return v; }
// two consecutive samples from global memory, any alignment (L1-allocating: neighbouring loads share sectors)
template <class S>
__device__ __forceinline__ float2 ldg_pair(const S* p) {
    if constexpr (sizeof(S) == 4) {
        if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return v; }
        return make_float2(ldg_elem(p), ldg_elem(p + 1));
    } else if constexpr (sizeof(S) == 2) {
        if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) { const short2 v = __ldg(reinterpret_cast<const short2*>(p)); return make_float2((float)v.x, (float)v.y); }
        return make_float2((float)ldg_elem(p), (float)ldg_elem(p + 1));
    } else {
        if ((reinterpret_cast<uintptr_t>(p) & 1) == 0) { const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(p)); return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
        return make_float2((float)(ldg_elem(p) - 128), (float)(ldg_elem(p + 1) - 128));
    }
}
// the same from the shared-memory staging buffer (byte address)
template <class S>
__device__ __forceinline__ float2 lds_pair(const unsigned char* p, bool aligned) {
    if constexpr (sizeof(S) == 4) {
        if (aligned) return *reinterpret_cast<const float2*>(p);
        return make_float2(lds_f32(p), lds_f32(p + 4));
    } else if constexpr (sizeof(S) == 2) {
        if (aligned) { const short2 v = *reinterpret_cast<const short2*>(p); return make_float2((float)v.x, (float)v.y); }
        return make_float2((float)lds_s16(p), (float)lds_s16(p + 2));
    } else {
        if (aligned) { const uchar2 v = *reinterpret_cast<const uchar2*>(p); return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
        return make_float2((float)(lds_u8(p) - 128), (float)(lds_u8(p + 1) - 128));
    }
}

// ---- staged input (cp.async.bulk + mbarrier): one thread asks the copy engine for the next batch of symbols --
// each as the 16-byte aligned superset of its (arbitrarily aligned) samples -- while the CTA computes on the
// current one; the FFT's first pass then reads shared memory.  This is how narrow PCM samples enter the kernels
// (no float copy of the recording in HBM) and how raw-stream packets at odd sample offsets keep full-width
// global accesses; for aligned float32 batches it is a measured alternative to direct loads (GF3_RX_STAGED=1).
template <class P, class S> __host__ __device__ constexpr int raw_sym_bytes() { return P::N * (int)sizeof(S) + 16; }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Load one symbol's N samples (CP already skipped by the caller) into the first-pass layout
// x[i] = z[t + i*T],  z[m] = (s[2m], s[2m+1]).
template <class P, class S, int I0 = 0, int I1 = P::R>
__device__ __forceinline__ void load_symbol(float2 (&x)[P::R], const S* __restrict__ s, int t) {
    if constexpr (sizeof(S) == 4) {
        if ((reinterpret_cast<uintptr_t>(s) & 7) == 0) {
#pragma unroll
            for (int i = I0; i < I1; ++i) x[i] = ldg_stream2(reinterpret_cast<const float*>(s) + 2 * (t + i * P::T));
            return;
        }
    }
    // odd sample offset (arbitrary sync index) or PCM: scalar / narrow loads.  Each touches part of the same
    // sectors as its neighbours, so these loads DO allocate in L1.  (Measured alternative for float32: 8-byte aligned
    // loads of (s[2m-1], s[2m]) completed by a lane shuffle -- 32 SHFL per symbol and spills at 128 registers made the
    // odd-offset data symbols 2.2x SLOWER, 0.610 vs 0.273 ms per 1024 C3 streams; profiles/r02_summary.md)
#pragma unroll
    for (int i = I0; i < I1; ++i) x[i] = ldg_pair<S>(s + 2 * (t + i * P::T));
}

// predicated one-byte shared-memory store: the address is an operand, so the compiler cannot sink its
// computation into a branch around the store
__device__ __forceinline__ void sts_u8_if(unsigned addr, unsigned val, unsigned cond) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t}" ::"r"(addr), "r"(val), "r"(cond) : "memory");
}
template <class P, int NT>
__host__ __device__ constexpr size_t demod_stage_offset() {    // shared memory: [zbuf | tw | (16-byte aligned) stage | xorw]
    return (((size_t)((NT / P::T) * P::MP + P::TW_TOTAL) * sizeof(float2)) + 15) & ~(size_t)15;
}
// pilot blocks the fused estimate processes at a time: both when their sums and spectra fit the
// data-symbol kernel's spectrum buffer, else one after the other (N = 4096)
template <class P, int NT>
__host__ __device__ constexpr int demod_est_par() {
    return (2 * (P::M + P::MP) <= (NT / P::T) * P::MP && 2 * P::T <= NT) ? 2 : 1;
}
constexpr int GF3_FUSE_UNFIT = 1;      // internal: the fused estimate does not fit this geometry, use two launches
constexpr int kReseed = 64;     // data symbols per work item = distance between exact re-seeds of the equaliser recurrence

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
// Exact phases for a SHORT fit window.  The slope of OFDM.py:462 is a least-squares fit through the phases of
// He / Hs inside the window; the equaliser multiplies it by up to n * w ~ K.  With hundreds of bins in the
// window the float32 phase error of the channel estimate averages out (slope error ~1e-9), but a window of a
// few bins -- the reference's literal [500:1000] clips to bins 500..510 at K = 511 -- passes it on almost
// undamped: 3e-6 rad of phase error becomes 5e-4 rad at the band edge.  For windows of at most
// kFitExactMax bins the phases are therefore recomputed in double precision: double sums of the P known
// symbols, then a direct DFT of the window's bins (a rotation recurrence per lane, one warp per bin).
// sd: N doubles of shared scratch.  Called by all NT threads; phi is complete after the trailing barrier.
// ------------------------------------------------------------------------------------------
constexpr int kFitExactMax = 32;
template <class P, int NT, class S>
__device__ __noinline__ void refine_fit_phases(const S* base0, const S* base1, int symlen, int Pn, const float2* __restrict__ known,
                                               int flo, int nfit, double* sd, double* phi, int phi_blk_stride, int phi_off) {
    constexpr int N = P::N, NW = NT / 32, C = N / 32;       // samples per lane
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int blk = 0; blk < 2; ++blk) {
        const S* b = blk ? base1 : base0;
        __syncthreads();                                   // sd is free
        for (int j = tid; j < N; j += NT) {
            double acc = 0.0;
            for (int p = 0; p < Pn; ++p) acc += (double)cvt_sample<S>((S)ldg_elem(b + (int64_t)p * symlen + j));
            sd[j] = acc;
        }
        __syncthreads();
        for (int i = warp; i < nfit; i += NW) {
            const int k = flo + i + 1;                     // FFT bin of carrier index flo + i
            const int j0 = lane * C;
            double wr, wi, cr, ci;
            sincospi(-2.0 * (double)k / (double)N, &wi, &wr);                              // step e^{-2 pi i k / N}
            sincospi(-2.0 * (double)((int64_t)k * j0 % N) / (double)N, &ci, &cr);          // e^{-2 pi i k j0 / N}
            double xr = 0.0, xi = 0.0;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const double v = sd[j0 + c];
                xr = fma(v, cr, xr);
                xi = fma(v, ci, xi);
                const double t = cr * wr - ci * wi;
                ci = fma(cr, wi, ci * wr);
                cr = t;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xr += __shfl_xor_sync(0xffffffffu, xr, o);
                xi += __shfl_xor_sync(0xffffffffu, xi, o);
            }
            if (lane == 0) {
                const float2 kn = known[k - 1];            // H = X / known / P: the phase of X conj(known)
                const double hr = xr * (double)kn.x + xi * (double)kn.y, hi = xi * (double)kn.x - xr * (double)kn.y;
                phi[blk * phi_blk_stride + phi_off + i] = atan2(hi, hr);
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Channel estimate of ONE packet by all NT threads of a CTA (OFDM.py:407-418,593,429-462): the
// same four steps as rx_estimate_kernel below, callable from inside the data-symbol kernel so that
// the whole receive chain of a packet is one launch.
//   1. time-domain sum of the P leading / trailing known symbols (the FFT is linear)
//   2. real FFT of the sum, 3. divide by the known symbol -> Hs / He (global) and the fit-window
//   phases (shared), 4. unwrap + least-squares slope.
// Scratch: work = float2[>= PAR * (M + MP)] where PAR (1 or 2) pilot blocks are processed at a time;
// phi = double[2 * (fit_hi - fit_lo)].  Contains __syncthreads; the slope is returned to every thread.
// ------------------------------------------------------------------------------------------
template <class P, int NT, int PAR, class S>
__device__ __noinline__ double estimate_packet(const S* pkt_base, int symlen, int cp, int Pn, int Ln,
                                               const float2* __restrict__ known, float2* __restrict__ Hs,
                                               float2* __restrict__ He, int fit_lo, int fit_hi, float2* work,
                                               double* phi, const float2* tw, int* warp_tot, double* red,
                                               double* s_slope) {
    constexpr int T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, K = M - 1;
    static_assert(PAR * T <= NT, "not enough threads for the pilot-block FFTs");
    const int tid = threadIdx.x;
    const int flo = max(0, min(fit_lo, K)), fhi = max(flo, min(fit_hi, K)), nfit = fhi - flo;
    float2* avg = work;                    // [PAR][M]
    float2* zb = work + PAR * M;           // [PAR][MP]
    const float invP = 0.5f / (float)Pn;   // the untangled values are 2X
#pragma unroll 1
    for (int b0 = 0; b0 < 2; b0 += PAR) {
        // ---- 1. sums (pure streaming: many 16-byte loads in flight)
        const S* base0 = pkt_base + cp;
        const S* base1 = pkt_base + (int64_t)(Pn + Ln) * symlen + cp;
        const bool al16 = sizeof(S) == 4 && ((reinterpret_cast<uintptr_t>(base0) | reinterpret_cast<uintptr_t>(base1)) & 15) == 0 && (symlen % 4 == 0);
        if (al16) {
            constexpr int U = GF3_EST_U;
            for (int q = tid; q < PAR * (N / 4); q += NT) {
                const int bl = q / (N / 4), c4 = q % (N / 4);
                const float* s0 = reinterpret_cast<const float*>((b0 + bl) ? base1 : base0) + 4 * c4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p0 = 0; p0 < Pn; p0 += U) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int p = p0 + u < Pn ? p0 + u : Pn - 1;
                        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(s0 + (int64_t)p * symlen));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (p0 + u < Pn) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
                }
                *reinterpret_cast<float4*>(&avg[bl * M + 2 * c4]) = acc;
            }
        } else {
            for (int col = tid; col < PAR * M; col += NT) {
                const int bl = col / M, m = col % M;
                const S* s0 = ((b0 + bl) ? base1 : base0) + 2 * m;
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
                for (int p = 0; p < Pn; ++p) {
                    const float2 v = ldg_pair<S>(s0 + (int64_t)p * symlen);
                    acc.x += v.x;
                    acc.y += v.y;
                }
                avg[col] = acc;
            }
        }
        __syncthreads();
        // ---- 2. FFT of the sums (warps holding an active symbol group take part as whole warps)
        constexpr int ACTIVE = (PAR * T < 32) ? 32 : PAR * T;
        if (tid < ACTIVE) {
            float2 x[R];
            const int g = (tid / T) % PAR, t = tid % T;
#pragma unroll
            for (int i = 0; i < R; ++i) x[i] = avg[g * M + t + i * T];
            fft_forward<P, (T == ACTIVE && ACTIVE != NT) ? NT : ACTIVE>(x, zb + g * MP, tw, t, g);
        }
        __syncthreads();
        // ---- 3. untangle, divide by the known symbol, Hs / He to global, fit-window phases to smem
        for (int item = tid; item < PAR * (M / 2); item += NT) {
            const int bl = item / (M / 2), j = item % (M / 2);
            const int blk = b0 + bl;
            const int k = j == 0 ? M / 2 : j, km = M - k;
            const float2* zs = zb + bl * MP;
            float sn, cs;
            sincospif(2.0f * (float)k / (float)N, &sn, &cs);
            const float2 w2 = make_float2(-sn, -cs);
            const float2 z1 = zs[zpad<P>(k)], z2 = zs[zpad<P>(km)];
            const float2 s = make_float2(z1.x + z2.x, z1.y - z2.y);
            const float2 d = make_float2(z1.x - z2.x, z1.y + z2.y);
            const float2 tt = cmul(w2, d);
            const float2 x1 = cadd(s, tt);
            const float2 x2 = make_float2(s.x - tt.x, tt.y - s.y);
            float2* Hout = blk ? He : Hs;
            {
                const float2 kn = known[k - 1];           // |known| = 1: 1/known = conj(known)
                float2 h = cmul(x1, cconj(kn));
                h.x *= invP; h.y *= invP;
                Hout[k - 1] = h;
                if (k - 1 >= flo && k - 1 < fhi) phi[blk * nfit + (k - 1 - flo)] = atan2((double)h.y, (double)h.x);
            }
            if (j != 0) {
                const float2 kn = known[km - 1];
                float2 h = cmul(x2, cconj(kn));
                h.x *= invP; h.y *= invP;
                Hout[km - 1] = h;
                if (km - 1 >= flo && km - 1 < fhi) phi[blk * nfit + (km - 1 - flo)] = atan2((double)h.y, (double)h.x);
            }
        }
        __syncthreads();
    }
    // ---- 4. slope (short windows: phases recomputed in double precision first)
    if (nfit >= 2 && nfit <= kFitExactMax)
        refine_fit_phases<P, NT, S>(pkt_base + cp, pkt_base + (int64_t)(Pn + Ln) * symlen + cp, symlen, Pn, known, flo, nfit,
                                    reinterpret_cast<double*>(work), phi, nfit, 0);
    const double sl = fit_slope<NT>(phi - flo, nfit, flo, fhi, warp_tot, red);
    if (tid == 0) *s_slope = sl;
    __syncthreads();
    return *s_slope;
}

// ------------------------------------------------------------------------------------------
// Data-symbol kernel.  One CTA owns a run of 16-symbol chunks of ONE packet.
//   phase A: SF symbols at a time, T threads per symbol, FFT in registers -> Z in smem
//   phase B: thread <-> bin pair (k, M-k): real-FFT untangling, equaliser, demap -> 2-bit codes
//   flush  : 16 codes -> one 32-bit word of MSB-first packed bits, coalesced store
// ------------------------------------------------------------------------------------------
template <class P, int NT, int MINB, bool KNOWN_CH, bool WANT_EQ, bool FUSE_EST = false, class S = float, bool STAGED = false>
__global__ void __launch_bounds__(NT, MINB) rx_demod_kernel(const RxArgs a) {
    static_assert(!(KNOWN_CH && FUSE_EST), "the known-channel receiver has no estimate to fuse");
    static_assert(STAGED || sizeof(S) == 4, "PCM samples enter through the staged path");
    constexpr int T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP;
    constexpr int SF = NT / T;                        // symbols per FFT batch
    // symbols per packed-bit flush: a multiple of SF that divides kReseed, with FLUSH*Nd % 16 == 0, so
    // every chunk starts on a 32-bit word of the packet's bit stream (chosen by the launcher)
    const int FLUSH = a.flush;
    const int BATCHES = FLUSH / SF;
    constexpr int TB = (M / 2 < NT) ? M / 2 : NT;     // threads per symbol in phase B
    constexpr int SB = NT / TB;                       // symbols handled concurrently in phase B
    constexpr int PP = (M / 2) / TB;                  // bin pairs per thread
    constexpr int K = M - 1;
    // multi-warp symbol groups (N = 4096) have few loads per thread in flight and only two CTAs per
    // SM: issue the next batch's loads before the equaliser phase.  Half-warp groups (R = 32) have
    // no registers to spare for that and several CTAs per SM already overlap.
    constexpr int PREFETCH = STAGED ? 0 : (GF3_PREFETCH >= 0) ? GF3_PREFETCH : (T >= 128 ? 1 : 0);
    constexpr bool NAT = GF3_DEMOD_NATURAL != 0;      // last FFT pass leaves the spectrum unpadded (see fft_pass)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + SF * MP;
    uint8_t* stage = smem_raw + demod_stage_offset<P, NT>();                 // [FLUSH*Nd] 2-bit codes, one per byte

    const int tid = threadIdx.x;
    const int Nd = a.hi - a.lo;
    const int L = a.L;
    const int symlen = N + a.cp;
    const bool use_xor = a.xor2 != nullptr;
    const bool want_bits = a.bits != nullptr;
    const int stage_bytes = ((FLUSH * Nd + 15) & ~15) + 16;
    uint32_t* xorw = reinterpret_cast<uint32_t*>(stage + stage_bytes);       // [FLUSH*Nd/16 + 1]
    // staged input: SF raw symbols (16-byte aligned supersets) + the mbarrier their bulk copies complete on
    constexpr int RAWB = raw_sym_bytes<P, S>();
    [[maybe_unused]] unsigned char* raw = nullptr;
    [[maybe_unused]] uint64_t* mbar = nullptr;
    [[maybe_unused]] unsigned raw_parity = 0;
    if constexpr (STAGED) {
        const size_t off = (demod_stage_offset<P, NT>() + stage_bytes + (((size_t)FLUSH * Nd + 15) / 16 + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
        raw = smem_raw + off;
        mbar = reinterpret_cast<uint64_t*>(raw + (size_t)SF * RAWB);
        if (tid == 0) mbar_init(mbar, 1);
    }

    // ================= once per CTA (the CTA is persistent: it walks a contiguous range of blocks)
    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];
    // XOR decode (OFDM.py:541-544) is applied per packed 32-bit word at flush time: word w of a
    // chunk covers codes 16w..16w+15, i.e. data carriers (16w+i) mod Nd -- the same for every chunk
    if (use_xor && want_bits) {
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
    // ---- phase B identity.  A thread owns PP bin pairs (k, M-k), k = jb + pp*TB (k = M/2 takes the
    // slot of k = 0), and carries the two bins of a pair in the two lanes of FFMA2 / FADD2 / FMUL2.
    const int jb = tid % TB, sb = tid / TB;
    pk64 wA[PP], wB[PP];                                // w = -j e^{-2 pi i k / N} as A = (wr, wi), B = (-wi, wr)
    pk64 Gre[PP], Gim[PP];                              // rotating equaliser taps: lane x = bin k, lane y = bin M-k
    pk64 Ure[PP], Uim[PP];                              // per-symbol rotation: e^{-j delta} (exact) or Ure = tan(delta) (fast)
    pk64* const my_slot = reinterpret_cast<pk64*>(zbuf) + tid;   // scratch for p_opaque (zbuf is not in use yet)
    const pk64 pm = p_opaque(pk_pack(make_float2(1.f, -1.f)), my_slot);
    const pk64 tie_eps = p_opaque(pk_pack(make_float2(__uint_as_float(0x0D800000u), __uint_as_float(0x0D800000u))), my_slot);   // 2^-100
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
        const int j = jb + pp * TB;
        const int k = j == 0 ? M / 2 : j;
        float s, c;
        sincospif(2.0f * (float)k / (float)N, &s, &c);
        wA[pp] = p_opaque(pk_pack(make_float2(-s, -c)), my_slot);   // -j * exp(-2 pi i k / N) = -s - j c
        wB[pp] = p_opaque(pk_pack(make_float2(c, -s)), my_slot);
        Ure[pp] = Uim[pp] = Gre[pp] = Gim[pp] = 0ull;
    }
    // which of this thread's bins carry data: bit 2pp = bin k, bit 2pp+1 = bin M-k (absent for k = M/2).
    // Derived once per CTA and pinned in a register: recomputing it costs ~10 instructions per pair.
    // (A plan that is out of registers -- N = 2048 on 128-thread CTAs: 4 pairs x 5 packed constants next
    // to 32 complex samples -- rebuilds the mask per batch instead: cheaper than one more spill.)
    constexpr bool MASK_ONCE = GF3_DMASK_ONCE >= 0 ? GF3_DMASK_ONCE != 0 : !(R * PP >= 128 && MINB * NT >= 512);
    auto data_mask = [&]() {
        unsigned m = 0;
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int j = jb + pp * TB;
            const int k = j == 0 ? M / 2 : j, km = M - k;
            if (k >= a.lo && k < a.hi) m |= 1u << (2 * pp);
            if (j != 0 && km >= a.lo && km < a.hi) m |= 2u << (2 * pp);
        }
        asm volatile("" : "+r"(m));                     // pinned: never re-derived per store
        return m;
    };
    [[maybe_unused]] unsigned dmask_cta = 0;
    if constexpr (MASK_ONCE) dmask_cta = data_mask();
    const double inv_lp = 1.0 / (double)(L + a.P);
    __shared__ int est_warp_tot[NT / 32];
    __shared__ double est_red[NT / 32];
    __shared__ double est_slope;
    __syncthreads();

    // phase-A identity of this thread: symbol group ga, lane ta inside the group
    const int ga = tid / T, ta = tid % T;
    float2 x[R];

    // ================= work loop.  Work item = one flush chunk: FLUSH consecutive data symbols of one
    // packet (the last chunk of a packet may be shorter).  Each CTA takes a contiguous, equally sized
    // range of chunks.  The equaliser recurrence is seeded exactly (fp64 phase) at every multiple of
    // kReseed symbols; a CTA whose range starts between two such points seeds at the previous one and
    // replays the recurrence up to its first symbol with the very same FMAs, so no output bit depends
    // on how the chunks are spread over CTAs.
    const int cpp = a.chunks_per_packet;
    const int64_t total_chunks = a.n_packets * cpp;
    int64_t c_begin = total_chunks * blockIdx.x / gridDim.x;
    int64_t c_end = total_chunks * (blockIdx.x + 1) / gridDim.x;
#if GF3_WHOLE_PACKETS
    // experiment (off): ranges on packet boundaries, so that no packet's estimate is computed by two CTAs.  The chunk
    // ranges above stagger the CTAs' positions inside their packets, which decorrelates the load / FFT / estimate phases
    // of the CTAs sharing an SM; that is worth more than the 592 estimates saved
    if (FUSE_EST && a.n_packets >= 4 * (int64_t)gridDim.x) {
        c_begin = (a.n_packets * blockIdx.x / gridDim.x) * cpp;
        c_end = (a.n_packets * (blockIdx.x + 1) / gridDim.x) * cpp;
    }
#endif
    int64_t cur_pkt = -1;
    const S* const samples = reinterpret_cast<const S*>(a.samples);
    const S* pkt_base = nullptr;
    const float2* Hs = a.Hs;
    double slope = 0.0;
    bool fast_rot = false;

    auto sym_ptr_g = [&](int64_t c, int l_first, int g) -> const S* {
        // samples (after the cyclic prefix) of symbol g of the batch starting at symbol l_first of chunk c's
        // packet.  Symbols past the end of the packet are clamped to the last one: their spectra are computed
        // but never used, and no zero-fill is needed
        const int64_t pk = c / cpp;
        const S* base = pk == cur_pkt ? pkt_base : samples + (a.pkt_offset ? a.pkt_offset[pk] : pk * a.pkt_stride);
        int l = l_first + g;
        l = l < L ? l : L - 1;
        return base + (int64_t)(a.P + l) * symlen + a.cp;
    };
    auto sym_ptr = [&](int64_t c, int l_first) -> const S* { return sym_ptr_g(c, l_first, ga); };
    // staged input: thread 0 hands the batch's SF symbols to the copy engine (one bulk copy per symbol; the
    // cyclic prefix between them is never fetched)
    [[maybe_unused]] auto stage_issue = [&](int64_t c, int l_first) {
        unsigned total = 0;
        for (int g = 0; g < SF; ++g) {
            const uintptr_t A = reinterpret_cast<uintptr_t>(sym_ptr_g(c, l_first, g));
            total += (unsigned)(((A & 15) + (uintptr_t)N * sizeof(S) + 15) & ~(uintptr_t)15);
        }
        mbar_expect_tx(mbar, total);
        for (int g = 0; g < SF; ++g) {
            const uintptr_t A = reinterpret_cast<uintptr_t>(sym_ptr_g(c, l_first, g));
            const unsigned bytes = (unsigned)(((A & 15) + (uintptr_t)N * sizeof(S) + 15) & ~(uintptr_t)15);
            bulk_g2s(raw + (size_t)g * RAWB, reinterpret_cast<const void*>(A & ~(uintptr_t)15), bytes, mbar);
        }
    };
    // PREFETCH == 3: the first GF3_PREFETCH_PART registers of the next batch are requested before the equaliser phase
    // (a partial register prefetch: what fits next to the phase's own registers), the rest at the start of the FFT
    constexpr int PF = (PREFETCH == 3) ? (GF3_PREFETCH_PART < R ? GF3_PREFETCH_PART : R) : 0;
    if constexpr (PREFETCH == 1) {
        if (c_begin < c_end) load_symbol<P, S>(x, sym_ptr(c_begin, (int)(c_begin % cpp) * FLUSH), ta);
    }
    if constexpr (PREFETCH == 3) {
        if (c_begin < c_end) load_symbol<P, S, 0, PF>(x, sym_ptr(c_begin, (int)(c_begin % cpp) * FLUSH), ta);
    }
    if constexpr (STAGED) {       // (the barrier above made the mbarrier's initialisation visible)
        if (tid == 0 && c_begin < c_end) stage_issue(c_begin, (int)(c_begin % cpp) * FLUSH);
    }

#pragma unroll 1
    for (int64_t c = c_begin; c < c_end; ++c) {
        const int64_t pkt = c / cpp;
        const int l0 = (int)(c - pkt * cpp) * FLUSH;
        const int nsym = min(FLUSH, L - l0);
        if (pkt != cur_pkt) {                           // ---- once per packet
            cur_pkt = pkt;
            pkt_base = samples + (a.pkt_offset ? a.pkt_offset[pkt] : pkt * a.pkt_stride);
            if constexpr (!KNOWN_CH) {
                Hs = a.Hs + pkt * K;
                if constexpr (FUSE_EST) {
                    // the packet's channel estimate, computed here (a packet split between two CTAs is
                    // estimated by both: same inputs, same arithmetic, same values written)
                    constexpr int PAR = demod_est_par<P, NT>();
                    __syncthreads();                    // zbuf / stage are free: every earlier chunk is flushed
                    slope = estimate_packet<P, NT, PAR, S>(pkt_base, symlen, a.cp, a.P, L, a.known, const_cast<float2*>(Hs),
                                                        const_cast<float2*>(a.He) + pkt * K, a.fit_lo, a.fit_hi, zbuf,
                                                        reinterpret_cast<double*>(stage), tw, est_warp_tot, est_red, &est_slope);
                    if (tid == 0) const_cast<double*>(a.slope)[pkt] = slope;
                } else {
                    slope = a.slope[pkt];
                }
                // For the bits only the sign of data * G matters, so the per-symbol rotation
                // G *= e^{-j delta} may be replaced by G *= (1 - j tan(delta)) = e^{-j delta} / cos(delta):
                // the same angle, two FMAs.  Valid while cos(delta) > 0 and the growth over one re-seed
                // period stays far inside fp32: |delta| <= 0.7 rad gives at most 1.31^64 = 3e7.  Anything
                // else (and the constellation output, which needs |G| = |Hs|) takes the exact rotation.
                fast_rot = !WANT_EQ && fabs(slope) * (double)(K - 1) * (double)SB * inv_lp <= 0.7;
#pragma unroll
                for (int pp = 0; pp < PP; ++pp) {
                    const int j = jb + pp * TB;
                    const int k = j == 0 ? M / 2 : j, km = M - k;
                    const float2 r1 = expmj(slope * inv_lp * (double)((k - 1) * SB));
                    const float2 r2 = expmj(slope * inv_lp * (double)((km - 1) * SB));
                    pk64 ur, ui = 0ull;
                    if (fast_rot) {
                        ur = pk_pack(make_float2(-r1.y / r1.x, -r2.y / r2.x));      // tan(delta)
                    } else {
                        ur = pk_pack(make_float2(r1.x, r2.x));
                        ui = pk_pack(make_float2(r1.y, r2.y));
                    }
                    Ure[pp] = ur;
                    Uim[pp] = ui;
                }
            }
        }
        if (c == c_begin || (l0 % kReseed) == 0) {
            // ---- seed the rotating equaliser taps (OFDM.py:466-478) for symbol l_seed + sb, then replay
            const int l_seed = l0 - l0 % kReseed;
#pragma unroll
            for (int pp = 0; pp < PP; ++pp) {
                const int j = jb + pp * TB;
                const int k = j == 0 ? M / 2 : j, km = M - k;
                const float2 h1 = Hs[k - 1], h2 = Hs[(j == 0 ? k : km) - 1];
                float2 g1 = h1, g2 = h2;
                if constexpr (!KNOWN_CH) {
                    const double wl = ((double)(l_seed + sb) + 0.5 * (double)a.P) * inv_lp;   // OFDM.py:471,474
                    g1 = cmul(cconj(h1), expmj(slope * (double)(k - 1) * wl));
                    g2 = cmul(cconj(h2), expmj(slope * (double)(km - 1) * wl));
                }
                Gre[pp] = pk_pack(make_float2(g1.x, g2.x));
                Gim[pp] = pk_pack(make_float2(g1.y, g2.y));
            }
            if constexpr (!KNOWN_CH) {
                const int steps = (l0 - l_seed) / SB;
                if (steps > 0) {
#pragma unroll
                    for (int pp = 0; pp < PP; ++pp) {
                        const pk64 ur = Ure[pp], ui = Uim[pp];
                        pk64 gre = Gre[pp], gim = Gim[pp];
#pragma unroll 1
                        for (int i = 0; i < steps; ++i) {
                            const pk64 gr = gre;
                            if (fast_rot) {
                                gre = p_fma(ur, gim, gr);
                                gim = p_fma(p_neg(ur), gr, gim);
                            } else {
                                gre = p_fma(p_neg(gim), ui, p_mul(gr, ur));
                                gim = p_fma(gr, ui, p_mul(gim, ur));
                            }
                        }
                        Gre[pp] = gre;
                        Gim[pp] = gim;
                    }
                }
            }
        }

        {
            const int nb_eff = min(BATCHES, (nsym + SF - 1) / SF);      // batches that hold at least one symbol of the packet
#pragma unroll 1
            for (int b = 0; b < nb_eff; ++b) {
                // ---------------- phase A: FFT of SF symbols
#if GF3_ABL & 4
                {
#pragma unroll
                    for (int i = 0; i < R; ++i) x[i] = make_float2(__int_as_float(0x3f800000 | ((tid + i + b) << 8)), __int_as_float(0x3f800000 | ((tid * 3 + i) << 7)));
                }
#else
                if constexpr (STAGED) {
                    // this batch's symbols have landed in the staging buffer (bulk copies issued one batch ago)
                    mbar_wait(mbar, raw_parity);
                    raw_parity ^= 1u;
                    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(sym_ptr(c, l0 + b * SF)) & 15);
                    const unsigned char* rb = raw + (size_t)ga * RAWB + sh;
                    if ((sh & (2 * sizeof(S) - 1)) == 0) {         // pairs are naturally aligned: one load per point
#pragma unroll
                        for (int i = 0; i < R; ++i) x[i] = lds_pair<S>(rb + (size_t)(2 * (ta + i * T)) * sizeof(S), true);
                    } else {                                        // (a real branch: as a select both forms would be issued)
#pragma unroll
                        for (int i = 0; i < R; ++i) x[i] = lds_pair<S>(rb + (size_t)(2 * (ta + i * T)) * sizeof(S), false);
                    }
                    __syncthreads();                               // every thread has its samples: the buffer is free again
                    if (tid == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (b + 1 < nb_eff) stage_issue(c, l0 + (b + 1) * SF);
                        else if (c + 1 < c_end) stage_issue(c + 1, (int)((c + 1) % cpp) * FLUSH);
                    }
                } else if constexpr (PREFETCH == 3) load_symbol<P, S, PF, R>(x, sym_ptr(c, l0 + b * SF), ta);
                else if constexpr (PREFETCH != 1) load_symbol<P, S>(x, sym_ptr(c, l0 + b * SF), ta);
#endif
#if GF3_ABL & 2
                {
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < R; ++i) acc = cadd(acc, x[i]);
                    zbuf[ga * MP + ta] = acc;
                }
#else
                fft_forward<P, NT, NAT>(x, zbuf + ga * MP, tw, ta, ga);
#endif
                __syncthreads();
                if constexpr (PREFETCH == 1) {
                    // prefetch the next batch's samples (possibly the first batch of the next chunk of this
                    // CTA's range): the loads fly while phase B computes
                    if (b + 1 < nb_eff) {
                        load_symbol<P, S>(x, sym_ptr(c, l0 + (b + 1) * SF), ta);
                    } else if (c + 1 < c_end) {
                        load_symbol<P, S>(x, sym_ptr(c + 1, (int)((c + 1) % cpp) * FLUSH), ta);
                    }
                } else if constexpr (PREFETCH == 3) {
                    if (b + 1 < nb_eff) {
                        load_symbol<P, S, 0, PF>(x, sym_ptr(c, l0 + (b + 1) * SF), ta);
                    } else if (c + 1 < c_end) {
                        load_symbol<P, S, 0, PF>(x, sym_ptr(c + 1, (int)((c + 1) % cpp) * FLUSH), ta);
                    }
                } else if constexpr (PREFETCH == 2) {
                    const int nl = l0 + (b + 1) * SF + ga;
                    if (ta == 0 && b + 1 < BATCHES && nl < L) {
                        const S* sp = pkt_base + (int64_t)(a.P + nl) * symlen + a.cp;
                        const uintptr_t lo16 = reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo16), "r"(N * 4 + 16) : "memory");
                    }
                }
                // ---------------- phase B: untangle, equalise, demap
                // thread <-> PP bin pairs; walks the batch's symbols sb, sb+SB, ...
                if (b == 0 && ((nsym * Nd) & 15) && tid < 16) stage[nsym * Nd + tid] = 0;   // a partial last word's missing codes are 0
                {
                    const int ls0 = b * SF + sb;                               // first symbol (inside the chunk) of this thread
                    int n_it = (nsym - ls0 + SB - 1) / SB;                      // valid symbols for this thread in this batch
                    n_it = n_it < 0 ? 0 : (n_it > SF / SB ? SF / SB : n_it);
                    auto phase_b = [&](auto fastc, auto bitsc) {
                        [[maybe_unused]] constexpr bool FAST = decltype(fastc)::value;
                        constexpr bool BITS = decltype(bitsc)::value;
                        const float2 *zp1[PP], *zp2[PP];
                        unsigned sp1[PP], sp2[PP];                              // shared-memory byte addresses of the code slots
                        bool e2[PP];                                            // bin M-k exists (k != M/2)
                        const unsigned st0 = (unsigned)__cvta_generic_to_shared(stage) + ls0 * Nd - a.lo;
#pragma unroll
                        for (int pp = 0; pp < PP; ++pp) {
                            const int j = jb + pp * TB;
                            const int k = j == 0 ? M / 2 : j, km = M - k;
                            zp1[pp] = zbuf + sb * MP + (NAT ? k : zpad<P>(k));
                            zp2[pp] = zbuf + sb * MP + (NAT ? km : zpad<P>(km));
                            sp1[pp] = st0 + k;
                            sp2[pp] = st0 + km;
                            e2[pp] = j != 0;
                        }
                        unsigned dmask;
                        if constexpr (MASK_ONCE) dmask = dmask_cta;
                        else dmask = data_mask();
                        [[maybe_unused]] float2* eqp = nullptr;
                        if constexpr (WANT_EQ) eqp = a.eq + ((int64_t)pkt * L + l0 + ls0) * K - 1;
                        const int st_step = SB * Nd;
                        int soff = 0;
                        constexpr int UNR = (SF / SB) < GF3_PHASEB_UNROLL ? (SF / SB) : GF3_PHASEB_UNROLL;
#pragma unroll UNR
                        for (int it = 0; it < n_it; ++it) {
#pragma unroll
                            for (int pp = 0; pp < PP; ++pp) {
                                const pk64 z1 = *reinterpret_cast<const pk64*>(zp1[pp] + it * (SB * MP));
                                const pk64 z2 = *reinterpret_cast<const pk64*>(zp2[pp] + it * (SB * MP));
                                // s = Z[k] + conj Z[M-k], d = Z[k] - conj Z[M-k]
                                const pk64 sa = p_add(z1, z2);                      // (s.x, d.y)
                                const pk64 sd = p_sub(z1, z2);                      // (d.x, s.y)
                                const pk64 tt = p_fma(p_bc(p_hi(sa)), wB[pp], p_mul(p_bc(p_lo(sd)), wA[pp]));   // w * d
                                // 2 X[k] = s + tt, 2 X[M-k] = conj(s - tt): real and imaginary parts of the pair
                                const pk64 xre = p_fma(p_bc(p_lo(tt)), pm, p_bc(p_lo(sa)));
                                const pk64 xim = p_fma(p_bc(p_hi(sd)), pm, p_bc(p_hi(tt)));
                                // (the inner products are a * b + (+0): a sum with +0 is never -0, so neither is y -- the
                                // sign bit of a component is then exactly "component < 0", as the reference's argmin needs)
                                const pk64 yre_ = p_fma(p_neg(xim), Gim[pp], p_fma(xre, Gre[pp], 0ull));
                                const pk64 yim_ = p_fma(xre, Gim[pp], p_fma(xim, Gre[pp], 0ull));
                                if constexpr (!KNOWN_CH) {
                                    const pk64 gr = Gre[pp];
                                    if constexpr (FAST) {                          // G *= 1 - j tan(delta)
                                        Gre[pp] = p_fma(Ure[pp], Gim[pp], gr);
                                        Gim[pp] = p_fma(p_neg(Ure[pp]), gr, Gim[pp]);
                                    } else {                                       // G *= e^{-j delta}
                                        Gre[pp] = p_fma(p_neg(Gim[pp]), Uim[pp], p_mul(gr, Ure[pp]));
                                        Gim[pp] = p_fma(gr, Uim[pp], p_mul(Gim[pp], Ure[pp]));
                                    }
                                }
                                const float2 yre = pk_unpack(yre_), yim = pk_unpack(yim_);
                                if constexpr (BITS) {
                                    // OFDM.py:484-500: argmin over [(0,0),(1,0),(1,1),(0,1)] keeps the FIRST minimum, i.e.
                                    //   b1 = real < 0,  b0 = imag < 0 or (imag == 0 and real < 0)   (first-minimum rule, pinned on exact ties by the tests).
                                    // tie = imag + real * 2^-100 has the sign of imag unless imag is exactly zero, where it takes
                                    // the sign of real (an underflow to -0 keeps the sign bit); y itself is never -0 (above).
                                    const float2 tie = pk_unpack(p_fma(yre_, tie_eps, yim_));
                                    sts_u8_if(sp1[pp] + soff, __funnelshift_l(__float_as_uint(yre.x), __float_as_uint(tie.x) >> 31, 1), dmask & (1u << (2 * pp)));
                                    sts_u8_if(sp2[pp] + soff, __funnelshift_l(__float_as_uint(yre.y), __float_as_uint(tie.y) >> 31, 1), dmask & (2u << (2 * pp)));
                                }
                                if constexpr (WANT_EQ) {
                                    const int k = (int)(sp1[pp] - st0), km = M - k;
                                    float sc1 = 0.5f, sc2 = 0.5f;
                                    if constexpr (!KNOWN_CH) {
                                        // |H| = |Hs| + (|He| - |Hs|) w  (OFDM.py:471); G carries conj(Hs) unnormalised
                                        const float eq_w = (float)(((double)(l0 + ls0 + it * SB) + 0.5 * (double)a.P) * inv_lp);
                                        const float2 hs1 = Hs[k - 1], he1 = a.He[pkt * K + k - 1];
                                        const float a1 = sqrtf(hs1.x * hs1.x + hs1.y * hs1.y);
                                        const float e1 = sqrtf(he1.x * he1.x + he1.y * he1.y);
                                        sc1 = 0.5f / (a1 * (a1 + (e1 - a1) * eq_w));
                                        if (e2[pp]) {
                                            const float2 hs2 = Hs[km - 1], he2 = a.He[pkt * K + km - 1];
                                            const float a2 = sqrtf(hs2.x * hs2.x + hs2.y * hs2.y);
                                            const float e2v = sqrtf(he2.x * he2.x + he2.y * he2.y);
                                            sc2 = 0.5f / (a2 * (a2 + (e2v - a2) * eq_w));
                                        }
                                    }
                                    eqp[k] = make_float2(yre.x * sc1, yim.x * sc1);
                                    if (e2[pp]) eqp[km] = make_float2(yre.y * sc2, yim.y * sc2);
                                }
                            }
                            soff += st_step;
                            if constexpr (WANT_EQ) eqp += (int64_t)SB * K;
                        }
                    };
#if !(GF3_ABL & 1)
                    if (want_bits) {
                        if (fast_rot) phase_b(std::true_type{}, std::true_type{});
                        else phase_b(std::false_type{}, std::true_type{});
                    } else {
                        phase_b(std::false_type{}, std::false_type{});
                    }
#endif
                }
                __syncthreads();
            }

            // ---------------- flush: 16 two-bit codes -> one 32-bit word (MSB-first bytes)
            // four codes c0..c3 (one per byte of u) -> (c0<<6 | c1<<4 | c2<<2 | c3) is the top byte of
            // u * 0x40100401 (no carries: every partial product lands on its own 2-bit field)
            if (want_bits) {
                const int ncodes = nsym * Nd;
                const int nwords = (ncodes + 15) >> 4;
                const int nfull = ncodes >> 4;                      // words made of 16 real codes
                uint32_t* out = reinterpret_cast<uint32_t*>(a.bits + pkt * a.bits_stride) + (int64_t)l0 * Nd / 16;
                auto pack16 = [&](int w) -> uint32_t {
                    const uint4 v = *reinterpret_cast<const uint4*>(stage + 16 * w);
                    const uint32_t t0 = v.x * 0x40100401u, t1 = v.y * 0x40100401u, t2 = v.z * 0x40100401u, t3 = v.w * 0x40100401u;
                    return __byte_perm(__byte_perm(t0, t1, 0x0073), __byte_perm(t2, t3, 0x0073), 0x5410);
                };
                constexpr int FU = GF3_FLUSH_UNROLL > 0 ? GF3_FLUSH_UNROLL : (P::LOGN <= 10 ? 4 : 2);   // re-swept on the final kernel (r02ay / r02az)
                if (use_xor) {
#pragma unroll FU
                    for (int w = tid; w < nfull; w += NT) out[w] = pack16(w) ^ xorw[w];
                } else {
#pragma unroll FU
                    for (int w = tid; w < nfull; w += NT) out[w] = pack16(w);
                }
                if (nwords > nfull && tid == 0) {                   // the chunk's last, partial word: pad bits stay zero
                    uint32_t word = pack16(nfull);
                    if (use_xor) {
                        const int r = ncodes - 16 * nfull;          // codes in this word (1..15)
                        const int fb = r >> 2, rm = r & 3;
                        const uint32_t m = (fb ? (0xFFFFFFFFu >> (32 - 8 * fb)) : 0u) | (rm ? (((0xFF00u >> (2 * rm)) & 0xFFu) << (8 * fb)) : 0u);
                        word ^= xorw[nfull] & m;
                    }
                    out[nfull] = word;
                }
                if (l0 + nsym >= L) {                               // last chunk of the packet: clear the row's pad words
                    const int stride_words = (int)(a.bits_stride / 4) - (int)((int64_t)l0 * Nd / 16);
                    for (int w = nwords + tid; w < stride_words; w += NT) out[w] = 0u;
                }
                // no barrier here: the next chunk's loads and FFT do not touch the code staging area, and
                // the barrier that ends its first FFT phase orders this flush's reads before new codes
            }
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
    const void* samples;
    const int64_t* pkt_offset;
    const float2* known;     // [K]
    float2* Hs;
    float2* He;
    double* slope;
    const float2* tw;
    int64_t pkt_stride;
    int cp, P, L, fit_lo, fit_hi;
};

template <class P, class S = float>
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
    const S* pkt_base = reinterpret_cast<const S*>(a.samples) + (a.pkt_offset ? a.pkt_offset[pkt] : pkt * a.pkt_stride);
    const int symlen = N + a.cp;

    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // ---- 1. time-domain sums of the pilot symbols (pure streaming: keep many 16-byte loads in flight)
    {
        const S* blk0 = pkt_base + a.cp;
        const S* blk1 = pkt_base + (int64_t)(a.P + a.L) * symlen + a.cp;
        const bool al16 = sizeof(S) == 4 && ((reinterpret_cast<uintptr_t>(blk0) | reinterpret_cast<uintptr_t>(blk1)) & 15) == 0 && (symlen % 4 == 0);
        if (al16) {
            constexpr int U = 10;
            for (int q = tid; q < 2 * (N / 4); q += NT) {          // float4 column q of block q / (N/4)
                const int blk = q / (N / 4), c4 = q % (N / 4);
                const float* s0 = reinterpret_cast<const float*>(blk ? blk1 : blk0) + 4 * c4;
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
                const S* s0 = (blk ? blk1 : blk0) + 2 * m;
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
                for (int p = 0; p < a.P; ++p) {
                    const float2 v = ldg_pair<S>(s0 + (int64_t)p * symlen);
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
    if (fhi - flo >= 2 && fhi - flo <= kFitExactMax)       // short window: exact phases (zbuf is dead: N doubles of scratch)
        refine_fit_phases<P, NT, S>(pkt_base + a.cp, pkt_base + (int64_t)(a.P + a.L) * symlen + a.cp, symlen, a.P, a.known, flo, fhi - flo,
                                    reinterpret_cast<double*>(zbuf), phi, K, flo);
    const double sl = fit_slope<NT>(phi, K, flo, fhi, warp_tot, red);
    if (tid == 0) a.slope[pkt] = sl;
}

// ------------------------------------------------------------------------------------------ launchers
#ifndef GF3_DEMOD_THREADS
#define GF3_DEMOD_THREADS 128
#endif
// Plan used by the data-symbol kernel for each symbol size, its CTA size and CTAs per SM.
#ifndef GF3_DEMOD_MINB
#define GF3_DEMOD_MINB (512 / GF3_DEMOD_THREADS)
#endif
template <int LOGN> struct DemodCfg { using Plan = FftPlan<LOGN>; static constexpr int NT = GF3_DEMOD_THREADS, MINB = GF3_DEMOD_MINB; };
// N = 4096: 128 threads per symbol (16 x 16 x 8, last pass on adjacent columns), two symbols per 256-thread CTA.
// (A warp-per-symbol 64 x 32 plan with ~255 registers / thread was measured slower: 8 warps per SM cannot hide latency.)
#ifndef GF3_DEMOD12_THREADS
#define GF3_DEMOD12_THREADS 256
#define GF3_DEMOD12_MINB 2
#endif
#if GF3_RX12_ALT == 3
template <> struct DemodCfg<12> { using Plan = FftPlan12P; static constexpr int NT = GF3_DEMOD12_THREADS, MINB = GF3_DEMOD12_MINB; };
#elif GF3_RX12_ALT == 2
template <> struct DemodCfg<12> { using Plan = FftPlan12C; static constexpr int NT = 128, MINB = 2; };
#elif GF3_RX12_ALT
template <> struct DemodCfg<12> { using Plan = FftPlan12B; static constexpr int NT = GF3_DEMOD12_THREADS, MINB = GF3_DEMOD12_MINB; };
#else
template <> struct DemodCfg<12> { using Plan = FftPlan<12>; static constexpr int NT = GF3_DEMOD12_THREADS, MINB = GF3_DEMOD12_MINB; };
#endif
// N = 2048: a warp per symbol (32 x 32).  128-thread CTAs would give every thread 4 bin pairs next to
// its 32 complex samples and spill; 256 threads (8 symbols per batch, 2 pairs per thread) fit:
// chain 1.118 vs 1.154 ms on 2048 streams.
#ifndef GF3_DEMOD11_THREADS
#define GF3_DEMOD11_THREADS 256
#define GF3_DEMOD11_MINB 2
#endif
template <> struct DemodCfg<11> { using Plan = FftPlan<11>; static constexpr int NT = GF3_DEMOD11_THREADS, MINB = GF3_DEMOD11_MINB; };

template <int LOGN, bool KNOWN_CH, bool WANT_EQ, bool FUSE_EST = false, class S = float, bool STAGED = false>
static int launch_demod(const gf3_plan* plan, RxArgs a, int64_t n_packets, cudaStream_t st) {
    using P = typename DemodCfg<LOGN>::Plan;
    // CTA size: one symbol group needs P::T threads; small CTAs (several per SM) decorrelate the
    // load / FFT / equalise phases of co-resident CTAs
    constexpr int NT = (P::T > DemodCfg<LOGN>::NT) ? P::T : DemodCfg<LOGN>::NT;
    // the staging buffer caps the CTAs per SM below the direct path's (N = 1024: 82 KB for float32 -> 2, 57 KB for
    // uint8 -> 3): give the register allocator the matching budget instead of spilling at the direct path's cap
    constexpr int MINB_STAGED = sizeof(S) == 4 ? 2 : 3;
    constexpr int MINB = (STAGED && DemodCfg<LOGN>::MINB > MINB_STAGED) ? MINB_STAGED : DemodCfg<LOGN>::MINB;
    constexpr int SF = NT / P::T;
    static_assert(kReseed % SF == 0, "FFT batch must divide the re-seed block");
    const int Nd = a.hi - a.lo;
    // smallest flush period: multiple of SF, FLUSH*Nd % 16 == 0, at least 8 symbols (amortise the
    // flush); always a power of two <= 16 or SF itself, so it divides the re-seed block
    int flush = SF;
    while ((flush * Nd) % 16 != 0 || flush < 8) flush += SF;
    GF3_REQUIRE(kReseed % flush == 0, "rx_demod: flush period %d does not divide the re-seed block", flush);
    a.flush = flush;
    a.tw = plan->d_tw_demod;
    a.n_packets = n_packets;
    a.chunks_per_packet = (a.L + flush - 1) / flush;
    size_t smem = demod_stage_offset<P, NT>() + (((size_t)flush * Nd + 15) & ~(size_t)15) + 16
                  + (((size_t)flush * Nd + 15) / 16 + 1) * sizeof(uint32_t);
    if (STAGED) smem = ((smem + 15) & ~(size_t)15) + (size_t)SF * raw_sym_bytes<P, S>() + 16;     // raw symbols + mbarrier
    auto kern = rx_demod_kernel<P, NT, MINB, KNOWN_CH, WANT_EQ, FUSE_EST, S, STAGED>;
    GF3_REQUIRE(smem <= 227 * 1024, "rx_demod: %zu bytes of shared memory needed (> 227 KB)", smem);
    if constexpr (FUSE_EST) {
        // sequential pilot blocks (N = 4096) make the in-kernel estimate slower than a separate launch
        if (demod_est_par<P, NT>() < 2 && !GF3_FUSE_SEQUENTIAL) return GF3_FUSE_UNFIT;
        if (LOGN == 12 && !GF3_FUSE12 && !GF3_FUSE_SEQUENTIAL) return GF3_FUSE_UNFIT;
        // the fused estimate keeps its fit-window phases (2 x window doubles) in the code staging area
        const int K = P::M - 1;
        const int flo = a.fit_lo < 0 ? 0 : (a.fit_lo > K ? K : a.fit_lo), fhi = a.fit_hi < flo ? flo : (a.fit_hi > K ? K : a.fit_hi);
        if ((size_t)2 * (fhi - flo) * sizeof(double) > ((((size_t)flush * Nd + 15) & ~(size_t)15) + 16)) return GF3_FUSE_UNFIT;
    }
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent CTAs: one full wave (resident CTAs per SM x SMs); each walks a contiguous, equally
    // sized range of flush chunks, so the per-CTA set-up (twiddles, XOR words, bin constants) is
    // paid once per CTA and not once per packet
    int per_sm = 0;
    GF3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total_blocks = n_packets * a.chunks_per_packet;
    int64_t grid = (int64_t)plan->sm_count * per_sm;
    if (grid > total_blocks) grid = total_blocks;
    if (const char* co = getenv("GF3_CARVEOUT")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
    if (getenv("GF3_DEBUG"))
        fprintf(stderr, "[gf3] rx_demod N=%d NT=%d smem=%zu B grid=%lld flush=%d blocks=%lld -> %d CTAs/SM%s (%d-byte samples)\n", P::N, NT, smem,
                (long long)grid, a.flush, (long long)total_blocks, per_sm, STAGED ? ", staged input (cp.async.bulk)" : "", (int)sizeof(S));
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

template <class P, class S = float>
static int launch_estimate(const gf3_plan* plan, EstArgs a, int64_t n_packets, cudaStream_t st) {
    const size_t smem = (size_t)(2 * P::M + 2 * P::MP + P::TW_TOTAL) * sizeof(float2);
    auto kern = rx_estimate_kernel<P, S>;
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


// Staged-input variants of the chain, one translation unit per sample type (gf3_rx_staged_*.cu):
//   fuse == 1: estimate + data symbols in one launch (returns GF3_FUSE_UNFIT if the geometry does not allow it)
//   fuse == 0: data symbols only (a.Hs / a.He / a.slope are inputs)
int rx_staged_demod_u8(const gf3_plan* plan, const RxArgs& a, int64_t n_packets, bool want_eq, bool fuse, cudaStream_t st);
int rx_staged_demod_i16(const gf3_plan* plan, const RxArgs& a, int64_t n_packets, bool want_eq, bool fuse, cudaStream_t st);
int rx_staged_demod_f32(const gf3_plan* plan, const RxArgs& a, int64_t n_packets, bool want_eq, bool fuse, cudaStream_t st);
int rx_estimate_u8(const gf3_plan* plan, const EstArgs& a, int64_t n_packets, cudaStream_t st);
int rx_estimate_i16(const gf3_plan* plan, const EstArgs& a, int64_t n_packets, cudaStream_t st);

}  // namespace gf3
