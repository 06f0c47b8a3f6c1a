// gf3_fft.cuh -- register-resident FFT engine shared by the RX, TX and sync kernels (sm_100a).
//
// The reference calls np.fft.fft / np.fft.ifft on N real-valued samples per OFDM symbol
// (OFDM.py:322,593).  Here a symbol's N real samples are packed as M = N/2 complex points
// z[m] = x[2m] + j x[2m+1]; a group of T threads holds them in registers (R points per thread)
// and runs a Stockham autosort FFT: in-register radix-R butterflies (twiddles folded to
// immediates at compile time) with one or two shared-memory exchanges.  No tensor cores: none
// of this is a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include <utility>

namespace gf3 {

// ---------------------------------------------------------------- compile-time loop helper
template <int I> using IC = std::integral_constant<int, I>;
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
    (f(IC<Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// ---------------------------------------------------------------- complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// cos(k*pi/16), k = 0..8 (exact to double precision, rounded to float at use)
__host__ __device__ constexpr double cos_pi16_q(int k) {
    return k == 0 ? 1.0
         : k == 1 ? 0.98078528040323044913
         : k == 2 ? 0.92387953251128675613
         : k == 3 ? 0.83146961230254523708
         : k == 4 ? 0.70710678118654752440
         : k == 5 ? 0.55557023301960222474
         : k == 6 ? 0.38268343236508977173
         : k == 7 ? 0.19509032201612826785
                  : 0.0;
}
// cos(i*pi/16) for any integer i
__host__ __device__ constexpr double cos_pi16(int i) {
    int k = ((i % 32) + 32) % 32;
    return k <= 8 ? cos_pi16_q(k) : k <= 16 ? -cos_pi16_q(16 - k) : k <= 24 ? -cos_pi16_q(k - 16) : cos_pi16_q(32 - k);
}
__host__ __device__ constexpr double sin_pi16(int i) { return cos_pi16(i - 8); }

// a * exp(-j * IDX * pi/16)   (IDX compile-time; multiples of 45 degrees need no general multiply)
template <int IDX>
__device__ __forceinline__ float2 mul_w32(float2 a) {
    constexpr int k = ((IDX % 32) + 32) % 32;
    if constexpr (k == 0) return a;
    else if constexpr (k == 8) return make_float2(a.y, -a.x);
    else if constexpr (k == 16) return make_float2(-a.x, -a.y);
    else if constexpr (k == 24) return make_float2(-a.y, a.x);
    else if constexpr (k % 8 == 4) {
        constexpr float h = 0.70710678118654752440f;
        // (c - js) with |c| = |s| = sqrt(1/2)
        constexpr float c = (k == 4 || k == 28) ? h : -h;
        constexpr float s = (k == 4 || k == 12) ? h : -h;
        return make_float2(c * a.x + s * a.y, c * a.y - s * a.x);
    } else {
        constexpr float c = (float)cos_pi16(k);
        constexpr float s = (float)sin_pi16(k);
        return make_float2(fmaf(c, a.x, s * a.y), fmaf(c, a.y, -s * a.x));
    }
}

// ---------------------------------------------------------------- in-register DFT codelets
// Forward DFT (e^{-2 pi i nk/R}), natural order in and out, on v[0..R).
template <int R> struct Dft;

template <> struct Dft<1> {
    static __device__ __forceinline__ void run(float2*) {}
};
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(float2* v) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    static __device__ __forceinline__ void run(float2* v) {
        float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        float2 s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
        float2 jd = make_float2(d13.y, -d13.x);   // -j * d13
        v[0] = cadd(s02, s13);
        v[1] = cadd(d02, jd);
        v[2] = csub(s02, s13);
        v[3] = csub(d02, jd);
    }
};
// Cooley-Tukey in registers: R = A*B, n = B*n1 + n2, k = k1 + A*k2.
template <int R, int A, int B>
__device__ __forceinline__ void dft_ct(float2* v) {
    static_assert(A * B == R && 32 % R == 0, "bad split");
    float2 y[B][A];
    static_for<B>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        float2 tmp[A];
        static_for<A>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; tmp[n1] = v[B * n1 + n2]; });
        Dft<A>::run(tmp);
        static_for<A>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            y[n2][k1] = mul_w32<n2 * k1 * (32 / R)>(tmp[k1]);
        });
    });
    static_for<A>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 tmp[B];
        static_for<B>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; tmp[n2] = y[n2][k1]; });
        Dft<B>::run(tmp);
        static_for<B>([&](auto k2c) { constexpr int k2 = decltype(k2c)::value; v[k1 + A * k2] = tmp[k2]; });
    });
}
template <> struct Dft<8> {
    static __device__ __forceinline__ void run(float2* v) { dft_ct<8, 2, 4>(v); }
};
template <> struct Dft<16> {
    static __device__ __forceinline__ void run(float2* v) { dft_ct<16, 4, 4>(v); }
};
template <> struct Dft<32> {
    static __device__ __forceinline__ void run(float2* v) { dft_ct<32, 4, 8>(v); }
};

// ---------------------------------------------------------------- FFT plans
// N real samples -> M = N/2 complex points; T threads per symbol, R points per thread,
// Stockham passes with radices RAD(0)=R, RAD(1), [RAD(2)].
template <int LOGN_, int R_, int NPASS_, int R0_, int R1_, int R2_>
struct FftPlanT {
    static constexpr int LOGN = LOGN_, N = 1 << LOGN_, M = N / 2, R = R_, T = M / R_;
    static constexpr int NPASS = NPASS_;
    static constexpr int LOGPAD = (R_ == 32 ? 5 : R_ == 16 ? 4 : 3);
    static constexpr int MP = M + (M >> LOGPAD);   // padded complex points per symbol
    __host__ __device__ static constexpr int rad(int p) { return p == 0 ? R0_ : p == 1 ? R1_ : R2_; }
    __host__ __device__ static constexpr int ns(int p) { return p == 0 ? 1 : p == 1 ? R0_ : R0_ * R1_; }
    // twiddle table: pass p >= 1 holds Q*(rad-1)*T entries at offset tw_off(p)
    static constexpr int TW1 = (R_ / R1_) * (R1_ - 1) * T;
    static constexpr int TW2 = NPASS_ > 2 ? (R_ / R2_) * (R2_ - 1) * T : 0;
    __host__ __device__ static constexpr int tw_off(int p) { return p <= 1 ? 0 : TW1; }
    static constexpr int TW_TOTAL = TW1 + TW2;
    static_assert(R0_ == R_ && R0_ * R1_ * R2_ == M, "radices must multiply to M and start with R");
};
template <int LOGN> struct FftPlan;
#define GF3_PLAN(LOGN_, R_, NPASS_, R0_, R1_, R2_) \
    template <> struct FftPlan<LOGN_> : FftPlanT<LOGN_, R_, NPASS_, R0_, R1_, R2_> {};
GF3_PLAN(6, 8, 2, 8, 4, 1)       // N=64    M=32    T=4
GF3_PLAN(7, 8, 2, 8, 8, 1)       // N=128   M=64    T=8
GF3_PLAN(8, 16, 2, 16, 8, 1)     // N=256   M=128   T=8
GF3_PLAN(9, 16, 2, 16, 16, 1)    // N=512   M=256   T=16
GF3_PLAN(10, 32, 2, 32, 16, 1)   // N=1024  M=512   T=16
GF3_PLAN(11, 32, 2, 32, 32, 1)   // N=2048  M=1024  T=32
GF3_PLAN(12, 16, 3, 16, 16, 8)   // N=4096  M=2048  T=128
#undef GF3_PLAN

template <class P>
__device__ __forceinline__ int zpad(int i) { return i + (i >> P::LOGPAD); }

// Synchronise the T threads that share one symbol.  `grp` = symbol-group index inside the CTA.
template <class P, int NTHREADS>
__device__ __forceinline__ void group_sync(int grp) {
    if constexpr (P::T <= 32) {
        __syncwarp();
    } else if constexpr (P::T == NTHREADS) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(P::T) : "memory");
    }
}

// One Stockham pass.  x[q*RAD + i]: thread-local data; zs: this symbol's padded smem buffer;
// tw: twiddle table in smem; t: thread index within the symbol group.
template <class P, int NTHREADS, int PASS>
__device__ __forceinline__ void fft_pass(float2 (&x)[P::R], float2* __restrict__ zs,
                                         const float2* __restrict__ tw, int t, int grp) {
    constexpr int RAD = P::rad(PASS), NS = P::ns(PASS), Q = P::R / RAD, STRIDE = P::M / RAD;
    if constexpr (PASS > 0) {
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const int j = t + q * P::T;
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                x[q * RAD + i] = zs[zpad<P>(j + i * STRIDE)];
            });
        });
        const float2* twp = tw + P::tw_off(PASS);
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            static_for<RAD - 1>([&](auto ic) {
                constexpr int i = decltype(ic)::value + 1;
                x[q * RAD + i] = cmul(x[q * RAD + i], twp[(q * (RAD - 1) + (i - 1)) * P::T + t]);
            });
        });
        group_sync<P, NTHREADS>(grp);   // every thread of the symbol has read before anyone overwrites
    }
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        Dft<RAD>::run(&x[q * RAD]);
    });
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const int j = t + q * P::T;
        const int base = (j / NS) * (NS * RAD) + (j % NS);
        static_for<RAD>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            zs[zpad<P>(base + i * NS)] = x[q * RAD + i];
        });
    });
    group_sync<P, NTHREADS>(grp);
}

// Full forward FFT of one symbol.  On entry x[i] = z[t + i*T] (i < R); on exit the natural-order
// spectrum Z[0..M) sits in zs (padded indexing) and is visible to the symbol's T threads.
template <class P, int NTHREADS>
__device__ __forceinline__ void fft_forward(float2 (&x)[P::R], float2* __restrict__ zs,
                                            const float2* __restrict__ tw, int t, int grp) {
    fft_pass<P, NTHREADS, 0>(x, zs, tw, t, grp);
    fft_pass<P, NTHREADS, 1>(x, zs, tw, t, grp);
    if constexpr (P::NPASS > 2) fft_pass<P, NTHREADS, 2>(x, zs, tw, t, grp);
}

// Host-side: fill the twiddle table of plan P (double precision, rounded to float).
template <class P>
inline void fill_twiddles(float2* out) {
    for (int pass = 1; pass < P::NPASS; ++pass) {
        const int RAD = P::rad(pass), NS = P::ns(pass), Q = P::R / RAD;
        float2* o = out + P::tw_off(pass);
        for (int q = 0; q < Q; ++q)
            for (int i = 1; i < RAD; ++i)
                for (int t = 0; t < P::T; ++t) {
                    const int j = t + q * P::T;
                    const double ang = -2.0 * 3.14159265358979323846 * (double)((j % NS) * i) / (double)(NS * RAD);
                    o[(q * (RAD - 1) + (i - 1)) * P::T + t] = make_float2((float)cos(ang), (float)sin(ang));
                }
    }
}

}  // namespace gf3
