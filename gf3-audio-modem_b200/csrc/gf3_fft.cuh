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
// packed f32x2 add / sub (FADD2): one issue slot for both components of a complex number
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk_pack(float2 a) { pk64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 pk_unpack(pk64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { pk64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b))); return pk_unpack(d); }
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) { pk64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b))); return pk_unpack(d); }
#ifndef GF3_SCALAR_CADD
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return pk_sub(a, b); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 mul_nj(float2 a) { return make_float2(a.y, -a.x); }   // a * (-j)

// packed f32x2 multiply / fused multiply-add (FMUL2 / FFMA2): one issue slot for two fp32 lanes.  On B200
// a scalar 3-register FFMA runs at about half rate while FFMA2 reaches the full 128 FMA/clk/SM
// (tools/microbench/f32x2.cu).  ptxas uses the scalar-broadcast operand form for (s, s) pairs.
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) { pk64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b))); return pk_unpack(d); }
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) { pk64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b)), "l"(pk_pack(c))); return pk_unpack(d); }

// 128-bit shared-memory load that the compiler may not split into (bank-conflicting) 32-bit loads
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

// f32x2 arithmetic on values that stay packed in 64-bit register pairs (no repacking per use)
__device__ __forceinline__ pk64 p_add(pk64 a, pk64 b) { pk64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk64 p_sub(pk64 a, pk64 b) { pk64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk64 p_mul(pk64 a, pk64 b) { pk64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk64 p_fma(pk64 a, pk64 b, pk64 c) { pk64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk64 p_neg(pk64 a) { const float2 v = pk_unpack(a); return pk_pack(make_float2(-v.x, -v.y)); }
__device__ __forceinline__ pk64 p_bc(float s) { return pk_pack(make_float2(s, s)); }
__device__ __forceinline__ float p_lo(pk64 a) { return pk_unpack(a).x; }
__device__ __forceinline__ float p_hi(pk64 a) { return pk_unpack(a).y; }
// Hide how a per-thread constant was derived, so that it is kept in its register pair instead of
// being rebuilt from a related value (negate + move, or immediates) at every use inside the hot
// loop.  A volatile round trip through the thread's own shared-memory slot is opaque to both
// compiler stages.
__device__ __forceinline__ pk64 p_opaque(pk64 v, void* slot) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(slot);
    asm volatile("st.volatile.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
    pk64 r;
    asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(r) : "r"(addr) : "memory");
    return r;
}
// Compile-time cos / sin of 2*pi*num/den (exact quadrant reduction on the integers, Taylor series
// on [0, pi/4]); evaluated in double, rounded to float where used, so in-register twiddles become
// instruction immediates.
__host__ __device__ constexpr double cx_taylor_sin(double x) {
    double term = x, sum = x;
    for (int k = 1; k < 12; ++k) { term *= -x * x / ((2 * k) * (2 * k + 1)); sum += term; }
    return sum;
}
__host__ __device__ constexpr double cx_taylor_cos(double x) {
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 12; ++k) { term *= -x * x / ((2 * k - 1) * (2 * k)); sum += term; }
    return sum;
}
// first-quadrant helper: angle = (r/den) * pi/2 with 0 <= r < den
__host__ __device__ constexpr double cx_q_cos(long long r, long long den) {
    const double hp = 1.57079632679489661923;
    return 2 * r <= den ? cx_taylor_cos(hp * (double)r / (double)den) : cx_taylor_sin(hp * (double)(den - r) / (double)den);
}
__host__ __device__ constexpr double cx_q_sin(long long r, long long den) {
    const double hp = 1.57079632679489661923;
    return 2 * r <= den ? cx_taylor_sin(hp * (double)r / (double)den) : cx_taylor_cos(hp * (double)(den - r) / (double)den);
}
__host__ __device__ constexpr double cx_cos2pi(long long num, long long den) {
    const long long n = ((num % den) + den) % den, q = (4 * n) / den, r = 4 * n - q * den;
    return q == 0 ? cx_q_cos(r, den) : q == 1 ? -cx_q_sin(r, den) : q == 2 ? -cx_q_cos(r, den) : cx_q_sin(r, den);
}
__host__ __device__ constexpr double cx_sin2pi(long long num, long long den) {
    const long long n = ((num % den) + den) % den, q = (4 * n) / den, r = 4 * n - q * den;
    return q == 0 ? cx_q_sin(r, den) : q == 1 ? cx_q_cos(r, den) : q == 2 ? -cx_q_sin(r, den) : -cx_q_cos(r, den);
}

// exp(-2*pi*j*n/64) in the two pair forms a packed complex multiply needs: A = (c, -s), B = (s, c)
// (x + jy)(c - js) = x*A + y*B.  Read through the constant bank into uniform registers.
__constant__ float4 gf3_w64ab[64] = {
    {1.000000000e+00f, 0.000000000e+00f, 0.000000000e+00f, 1.000000000e+00f},
    {9.951847267e-01f, -9.801714033e-02f, 9.801714033e-02f, 9.951847267e-01f},
    {9.807852804e-01f, -1.950903220e-01f, 1.950903220e-01f, 9.807852804e-01f},
    {9.569403357e-01f, -2.902846773e-01f, 2.902846773e-01f, 9.569403357e-01f},
    {9.238795325e-01f, -3.826834324e-01f, 3.826834324e-01f, 9.238795325e-01f},
    {8.819212643e-01f, -4.713967368e-01f, 4.713967368e-01f, 8.819212643e-01f},
    {8.314696123e-01f, -5.555702330e-01f, 5.555702330e-01f, 8.314696123e-01f},
    {7.730104534e-01f, -6.343932842e-01f, 6.343932842e-01f, 7.730104534e-01f},
    {7.071067812e-01f, -7.071067812e-01f, 7.071067812e-01f, 7.071067812e-01f},
    {6.343932842e-01f, -7.730104534e-01f, 7.730104534e-01f, 6.343932842e-01f},
    {5.555702330e-01f, -8.314696123e-01f, 8.314696123e-01f, 5.555702330e-01f},
    {4.713967368e-01f, -8.819212643e-01f, 8.819212643e-01f, 4.713967368e-01f},
    {3.826834324e-01f, -9.238795325e-01f, 9.238795325e-01f, 3.826834324e-01f},
    {2.902846773e-01f, -9.569403357e-01f, 9.569403357e-01f, 2.902846773e-01f},
    {1.950903220e-01f, -9.807852804e-01f, 9.807852804e-01f, 1.950903220e-01f},
    {9.801714033e-02f, -9.951847267e-01f, 9.951847267e-01f, 9.801714033e-02f},
    {0.000000000e+00f, -1.000000000e+00f, 1.000000000e+00f, 0.000000000e+00f},
    {-9.801714033e-02f, -9.951847267e-01f, 9.951847267e-01f, -9.801714033e-02f},
    {-1.950903220e-01f, -9.807852804e-01f, 9.807852804e-01f, -1.950903220e-01f},
    {-2.902846773e-01f, -9.569403357e-01f, 9.569403357e-01f, -2.902846773e-01f},
    {-3.826834324e-01f, -9.238795325e-01f, 9.238795325e-01f, -3.826834324e-01f},
    {-4.713967368e-01f, -8.819212643e-01f, 8.819212643e-01f, -4.713967368e-01f},
    {-5.555702330e-01f, -8.314696123e-01f, 8.314696123e-01f, -5.555702330e-01f},
    {-6.343932842e-01f, -7.730104534e-01f, 7.730104534e-01f, -6.343932842e-01f},
    {-7.071067812e-01f, -7.071067812e-01f, 7.071067812e-01f, -7.071067812e-01f},
    {-7.730104534e-01f, -6.343932842e-01f, 6.343932842e-01f, -7.730104534e-01f},
    {-8.314696123e-01f, -5.555702330e-01f, 5.555702330e-01f, -8.314696123e-01f},
    {-8.819212643e-01f, -4.713967368e-01f, 4.713967368e-01f, -8.819212643e-01f},
    {-9.238795325e-01f, -3.826834324e-01f, 3.826834324e-01f, -9.238795325e-01f},
    {-9.569403357e-01f, -2.902846773e-01f, 2.902846773e-01f, -9.569403357e-01f},
    {-9.807852804e-01f, -1.950903220e-01f, 1.950903220e-01f, -9.807852804e-01f},
    {-9.951847267e-01f, -9.801714033e-02f, 9.801714033e-02f, -9.951847267e-01f},
    {-1.000000000e+00f, 0.000000000e+00f, 0.000000000e+00f, -1.000000000e+00f},
    {-9.951847267e-01f, 9.801714033e-02f, -9.801714033e-02f, -9.951847267e-01f},
    {-9.807852804e-01f, 1.950903220e-01f, -1.950903220e-01f, -9.807852804e-01f},
    {-9.569403357e-01f, 2.902846773e-01f, -2.902846773e-01f, -9.569403357e-01f},
    {-9.238795325e-01f, 3.826834324e-01f, -3.826834324e-01f, -9.238795325e-01f},
    {-8.819212643e-01f, 4.713967368e-01f, -4.713967368e-01f, -8.819212643e-01f},
    {-8.314696123e-01f, 5.555702330e-01f, -5.555702330e-01f, -8.314696123e-01f},
    {-7.730104534e-01f, 6.343932842e-01f, -6.343932842e-01f, -7.730104534e-01f},
    {-7.071067812e-01f, 7.071067812e-01f, -7.071067812e-01f, -7.071067812e-01f},
    {-6.343932842e-01f, 7.730104534e-01f, -7.730104534e-01f, -6.343932842e-01f},
    {-5.555702330e-01f, 8.314696123e-01f, -8.314696123e-01f, -5.555702330e-01f},
    {-4.713967368e-01f, 8.819212643e-01f, -8.819212643e-01f, -4.713967368e-01f},
    {-3.826834324e-01f, 9.238795325e-01f, -9.238795325e-01f, -3.826834324e-01f},
    {-2.902846773e-01f, 9.569403357e-01f, -9.569403357e-01f, -2.902846773e-01f},
    {-1.950903220e-01f, 9.807852804e-01f, -9.807852804e-01f, -1.950903220e-01f},
    {-9.801714033e-02f, 9.951847267e-01f, -9.951847267e-01f, -9.801714033e-02f},
    {0.000000000e+00f, 1.000000000e+00f, -1.000000000e+00f, 0.000000000e+00f},
    {9.801714033e-02f, 9.951847267e-01f, -9.951847267e-01f, 9.801714033e-02f},
    {1.950903220e-01f, 9.807852804e-01f, -9.807852804e-01f, 1.950903220e-01f},
    {2.902846773e-01f, 9.569403357e-01f, -9.569403357e-01f, 2.902846773e-01f},
    {3.826834324e-01f, 9.238795325e-01f, -9.238795325e-01f, 3.826834324e-01f},
    {4.713967368e-01f, 8.819212643e-01f, -8.819212643e-01f, 4.713967368e-01f},
    {5.555702330e-01f, 8.314696123e-01f, -8.314696123e-01f, 5.555702330e-01f},
    {6.343932842e-01f, 7.730104534e-01f, -7.730104534e-01f, 6.343932842e-01f},
    {7.071067812e-01f, 7.071067812e-01f, -7.071067812e-01f, 7.071067812e-01f},
    {7.730104534e-01f, 6.343932842e-01f, -6.343932842e-01f, 7.730104534e-01f},
    {8.314696123e-01f, 5.555702330e-01f, -5.555702330e-01f, 8.314696123e-01f},
    {8.819212643e-01f, 4.713967368e-01f, -4.713967368e-01f, 8.819212643e-01f},
    {9.238795325e-01f, 3.826834324e-01f, -3.826834324e-01f, 9.238795325e-01f},
    {9.569403357e-01f, 2.902846773e-01f, -2.902846773e-01f, 9.569403357e-01f},
    {9.807852804e-01f, 1.950903220e-01f, -1.950903220e-01f, 9.807852804e-01f},
    {9.951847267e-01f, 9.801714033e-02f, -9.801714033e-02f, 9.951847267e-01f}};

// packed complex multiply: a * w with w given as A = (wr, wi), B = (-wi, wr); FMUL2 + FFMA2, the
// components of a enter through the scalar-broadcast operand form
__device__ __forceinline__ float2 cmul_ab(float2 a, float4 w) {
    return pk_fma(make_float2(a.y, a.y), make_float2(w.z, w.w), pk_mul(make_float2(a.x, a.x), make_float2(w.x, w.y)));
}

// a * exp(-2*pi*j * NUM/DEN)   (compile-time; multiples of 45 degrees need no general multiply)
template <int NUM, int DEN>
__device__ __forceinline__ float2 mul_w(float2 a) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    if constexpr (n == 0) return a;
    else if constexpr (4 * n == DEN) return make_float2(a.y, -a.x);        // -j
    else if constexpr (2 * n == DEN) return make_float2(-a.x, -a.y);       // -1
    else if constexpr (4 * n == 3 * DEN) return make_float2(-a.y, a.x);    // +j
    else {
#ifndef GF3_SCALAR_CMUL
        static_assert(64 % DEN == 0, "constant twiddle table covers divisors of 64");
        return cmul_ab(a, gf3_w64ab[n * (64 / DEN)]);
#else
        if constexpr ((8 * n) % DEN == 0) {
            constexpr float h = 0.70710678118654752440f;
            constexpr int o = (8 * n) / DEN;                                   // 1, 3, 5, 7
            constexpr float c = (o == 1 || o == 7) ? h : -h;                   // cos
            constexpr float sn = (o == 1 || o == 3) ? h : -h;                  // sin
            return make_float2(c * a.x + sn * a.y, c * a.y - sn * a.x);
        } else {
            constexpr float c = (float)cx_cos2pi(n, DEN);
            constexpr float sn = (float)cx_sin2pi(n, DEN);
            return make_float2(fmaf(c, a.x, sn * a.y), fmaf(c, a.y, -sn * a.x));
        }
#endif
    }
}

// ---------------------------------------------------------------- in-register DFT codelets
// Forward DFT (e^{-2 pi i nk/R}), natural order in and out, on v[0..R).
template <int R> struct Dft;

template <> struct Dft<1> {
    template <class V> static __device__ __forceinline__ void run(V*) {}
};
template <> struct Dft<2> {
    template <class V> static __device__ __forceinline__ void run(V* v) {
        V a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    template <class V> static __device__ __forceinline__ void run(V* v) {
        V s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        V s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
        V jd = mul_nj(d13);   // -j * d13
        v[0] = cadd(s02, s13);
        v[1] = cadd(d02, jd);
        v[2] = csub(s02, s13);
        v[3] = csub(d02, jd);
    }
};
// Cooley-Tukey in registers: R = A*B, n = B*n1 + n2, k = k1 + A*k2.
template <int R, int A, int B, class V>
__device__ __forceinline__ void dft_ct(V* v) {
    static_assert(A * B == R, "bad split");
    V y[B][A];
    static_for<B>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        V tmp[A];
        static_for<A>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; tmp[n1] = v[B * n1 + n2]; });
        Dft<A>::run(tmp);
        static_for<A>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            y[n2][k1] = mul_w<n2 * k1, R>(tmp[k1]);
        });
    });
    static_for<A>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        V tmp[B];
        static_for<B>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; tmp[n2] = y[n2][k1]; });
        Dft<B>::run(tmp);
        static_for<B>([&](auto k2c) { constexpr int k2 = decltype(k2c)::value; v[k1 + A * k2] = tmp[k2]; });
    });
}
template <> struct Dft<8> {
    template <class V> static __device__ __forceinline__ void run(V* v) { dft_ct<8, 2, 4, V>(v); }
};
template <> struct Dft<16> {
    template <class V> static __device__ __forceinline__ void run(V* v) { dft_ct<16, 4, 4, V>(v); }
};
template <> struct Dft<32> {
    template <class V> static __device__ __forceinline__ void run(V* v) { dft_ct<32, 4, 8, V>(v); }
};
template <> struct Dft<64> {
    template <class V> static __device__ __forceinline__ void run(V* v) { dft_ct<64, 8, 8, V>(v); }
};

// ---------------------------------------------------------------- FFT plans
// N real samples -> M = N/2 complex points; T threads per symbol, R points per thread,
// Stockham passes with radices RAD(0)=R, RAD(1), [RAD(2)].
#ifndef GF3_FFT_PAIRED
#define GF3_FFT_PAIRED 1
#endif
#ifndef GF3_TW_AB
#define GF3_TW_AB 0      // measured slower on C3 (1.23 vs 0.98 ms: twice the twiddle bytes through shared memory, spills at 128 registers)
#endif
template <int LOGN_, int R_, int NPASS_, int R0_, int R1_, int R2_, bool PAIRLAST_ = false>
struct FftPlanT {
    static constexpr int LOGN = LOGN_, N = 1 << LOGN_, M = N / 2, R = R_, T = M / R_;
    static constexpr int NPASS = NPASS_;
    static constexpr int LOGPAD = (R_ == 64 ? 6 : R_ == 32 ? 5 : R_ == 16 ? 4 : 3);
    // PAIRED (N = 1024: two passes, two sub-transforms per thread in the second): the second pass takes the
    // ADJACENT columns j = 2t, 2t+1 of the T x R0 intermediate matrix, so every shared-memory access of
    // the exchange is 128 bits wide (row stride R0 + 2 keeps 16-byte alignment and quarter-warp
    // conflict freedom): half the LDS / STS instructions of the FFT.
    static constexpr bool PAIRED = GF3_FFT_PAIRED && (LOGN_ == 10);
    static constexpr int ROWP = R0_ + 2;             // PAIRED: padded row of the intermediate matrix
    static constexpr int MP = PAIRED ? T * ROWP : M + (M >> LOGPAD);   // complex points reserved per symbol
    __host__ __device__ static constexpr int rad(int p) { return p == 0 ? R0_ : p == 1 ? R1_ : R2_; }
    __host__ __device__ static constexpr int ns(int p) { return p == 0 ? 1 : p == 1 ? R0_ : R0_ * R1_; }
    // twiddle table: pass p >= 1 holds Q*(rad-1)*T entries at offset tw_off(p).  DEDUP1: when T is a multiple of
    // the first radix, the pass-1 twiddle of column j = t + q T depends on j mod R0 = t mod R0 only, so the
    // table keeps (rad-1)*R0 entries (N = 4096: 240 instead of 1920 -- 13 KB of shared memory per CTA);
    // lanes t and t + R0 read the same address (broadcast), a warp touches one 128-byte line per load.
    static constexpr bool DEDUP1 = !(GF3_FFT_PAIRED && (LOGN_ == 10)) && (T > R0_) && (T % R0_ == 0);
    // TWAB (paired plan): every twiddle is stored as the two operand pairs of a packed complex multiply, A = (wr, wi) and
    // B = (-wi, wr), so x * w is FMUL2 + FFMA2 instead of 2 FMUL + 2 FFMA (twice the table, half the multiply instructions)
    static constexpr bool TWAB = GF3_TW_AB && PAIRED;
    // PAIRLAST (three-pass plans with two sub-transforms per thread in the last pass): the last pass takes the adjacent
    // columns j = 2t, 2t+1 as in PAIRED -- its loads, its twiddles and the natural-order stores are 128 bits wide; the
    // pass before it stores WITHOUT padding (its stores are runs of R0 consecutive points, conflict-free in any layout)
    static constexpr bool PAIRLAST = PAIRLAST_;
    static_assert(!PAIRLAST_ || (NPASS_ == 3 && R_ / R2_ == 2 && !PAIRED), "PAIRLAST: three passes, two sub-transforms in the last");
    static constexpr int TW1 = DEDUP1 ? (R1_ - 1) * R0_ : (TWAB ? 2 : 1) * (R_ / R1_) * (R1_ - 1) * T;
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
// N = 4096 as 32 x 8 x 8 on 64 threads (two warps per symbol, 32 points per thread): the data-symbol kernel's alternative
// to the 16 x 16 x 8 plan above (which the matched filter and the estimate kernel keep)
struct FftPlan12B : FftPlanT<12, 32, 3, 32, 8, 8> {};
// N = 4096 as 64 x 32 on ONE warp per symbol (64 points per thread, one exchange, __syncwarp only)
struct FftPlan12C : FftPlanT<12, 64, 2, 64, 32, 1> {};
// N = 4096 as 16 x 16 x 8 with the last pass on adjacent columns (128-bit loads, twiddles and stores)
struct FftPlan12P : FftPlanT<12, 16, 3, 16, 16, 8, true> {};

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

// Padded offset of a compile-time stride: for S a multiple of 2^LOGPAD,
//   zpad(b + i*S) = zpad(b) + i*(S + (S >> LOGPAD))          (no carry between the two terms),
// and for S == 1 with b a multiple of 2^LOGPAD and i < 2^LOGPAD, zpad(b + i) = zpad(b) + i.
// Folding this by hand keeps every shared-memory access of a pass at base + immediate.
template <class P, int S>
__host__ __device__ constexpr int zstride() {
    static_assert(S == 1 || S % (1 << P::LOGPAD) == 0, "stride must be 1 or a multiple of the padding period");
    return S == 1 ? 1 : S + (S >> P::LOGPAD);
}

// One Stockham pass.  x[q*RAD + i]: thread-local data; zs: this symbol's padded smem buffer;
// tw: twiddle table in smem; t: thread index within the symbol group.
//
// NATURAL (last pass only): the spectrum is written in natural order WITHOUT padding.  Every store
// instruction of the last pass writes runs of consecutive bins, which are conflict-free in any
// layout, and the bin-pair walk of the data-symbol kernel (ascending k, descending M-k) then reads
// conflict-free too (with padding, 16 descending bins straddle a pad slot and collide 2-way).
// STORE = false (last pass only): the pass ends with its results in registers -- x[q*RAD + i] is bin
// base(q) + i*NS, base(q) = (j / NS) * (NS * RAD) + j % NS, j = t + q*T (paired plans: j = 2t + q, i.e. x[i] and
// x[RAD + i] are the neighbouring bins 2t + i*NS and 2t + 1 + i*NS) -- for callers that consume the transform
// straight from registers (the matched filter and the transmitter write their output samples to global memory).
template <class P, int NTHREADS, int PASS, bool NATURAL = false, bool STORE = true>
__device__ __forceinline__ void fft_pass(float2 (&x)[P::R], float2* __restrict__ zs,
                                         const float2* __restrict__ tw, int t, int grp) {
    constexpr int RAD = P::rad(PASS), NS = P::ns(PASS), Q = P::R / RAD, STRIDE = P::M / RAD;
    static_assert(NS != 1 || RAD == (1 << P::LOGPAD), "first pass radix must equal the padding period");
    static_assert(!NATURAL || PASS == P::NPASS - 1, "natural-order output is for the last pass");
    if constexpr (PASS > 0) {
        if constexpr (P::PAIRED) {
            static_assert(Q == 2 && P::NPASS == 2, "paired layout: two passes, two sub-transforms in the second");
            const float4* src = reinterpret_cast<const float4*>(zs + 2 * t);
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const float4 v = src[i * (P::ROWP / 2)];                 // row i, columns 2t and 2t+1
                x[i] = make_float2(v.x, v.y);
                x[RAD + i] = make_float2(v.z, v.w);
            });
        } else if constexpr (P::PAIRLAST && PASS == P::NPASS - 1) {
            const float4* src = reinterpret_cast<const float4*>(zs + 2 * t);       // unpadded: columns 2t and 2t+1
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const float4 v = src[i * (STRIDE / 2)];
                x[i] = make_float2(v.x, v.y);
                x[RAD + i] = make_float2(v.z, v.w);
            });
        } else {
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float2* src = zs + zpad<P>(t + q * P::T);
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                x[q * RAD + i] = src[i * zstride<P, STRIDE>()];
            });
        });
        }
        if constexpr (P::PAIRED && P::TWAB) {
            // (A, B) pairs of the two adjacent columns: two 128-bit loads, two packed instructions per multiply
            // (layout [i][column parity][t]: the lanes of a load read consecutive 16-byte entries)
            const float4* twp = reinterpret_cast<const float4*>(tw + P::tw_off(PASS)) + t;
            static_for<RAD - 1>([&](auto ic) {
                constexpr int i = decltype(ic)::value + 1;
                x[i] = cmul_ab(x[i], twp[(i - 1) * 2 * P::T]);
                x[RAD + i] = cmul_ab(x[RAD + i], twp[(i - 1) * 2 * P::T + P::T]);
            });
        } else if constexpr (P::PAIRED || (P::PAIRLAST && PASS == P::NPASS - 1)) {
            // the twiddles of the two adjacent columns sit side by side: one 128-bit load for both
            const float4* twp = reinterpret_cast<const float4*>(tw + P::tw_off(PASS) + 2 * t);
            static_for<RAD - 1>([&](auto ic) {
                constexpr int i = decltype(ic)::value + 1;
                const float4 w = twp[(i - 1) * P::T];
                x[i] = cmul(x[i], make_float2(w.x, w.y));
                x[RAD + i] = cmul(x[RAD + i], make_float2(w.z, w.w));
            });
        } else if constexpr (PASS == 1 && P::DEDUP1) {
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float2* twp = tw + (t % NS);
            static_for<RAD - 1>([&](auto ic) {
                constexpr int i = decltype(ic)::value + 1;
                x[q * RAD + i] = cmul(x[q * RAD + i], twp[(i - 1) * NS]);
            });
        });
        } else {
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float2* twp = tw + P::tw_off(PASS) + t + q * ((RAD - 1) * P::T);
            static_for<RAD - 1>([&](auto ic) {
                constexpr int i = decltype(ic)::value + 1;
                x[q * RAD + i] = cmul(x[q * RAD + i], twp[(i - 1) * P::T]);
            });
        });
        }
        group_sync<P, NTHREADS>(grp);   // every thread of the symbol has read before anyone overwrites
    }
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        Dft<RAD>::run(&x[q * RAD]);
    });
    if constexpr (!STORE) {
        static_assert(PASS == P::NPASS - 1, "register output is for the last pass");
        return;
    } else
    if constexpr (P::PAIRED && PASS == 0) {
        // row t of the T x R0 matrix, two columns per 128-bit store
        float4* dst = reinterpret_cast<float4*>(zs + t * P::ROWP);
        static_for<RAD / 2>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            dst[m] = make_float4(x[2 * m].x, x[2 * m].y, x[2 * m + 1].x, x[2 * m + 1].y);
        });
    } else if constexpr ((P::PAIRED || P::PAIRLAST) && NATURAL) {
        // bins k = 2t + NS*i and k + 1 (the two sub-transforms of this thread) are neighbours
        float4* dst = reinterpret_cast<float4*>(zs + 2 * t);
        static_for<RAD>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            dst[i * (NS / 2)] = make_float4(x[i].x, x[i].y, x[RAD + i].x, x[RAD + i].y);
        });
    } else
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const int j = (P::PAIRED || (P::PAIRLAST && PASS == P::NPASS - 1)) ? 2 * t + q : t + q * P::T;
        const int base = (j / NS) * (NS * RAD) + (j % NS);
        if constexpr (NATURAL || (P::PAIRLAST && PASS == P::NPASS - 2)) {
            float2* dst = zs + base;
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                dst[i * NS] = x[q * RAD + i];
            });
        } else {
            float2* dst = zs + zpad<P>(base);
            static_for<RAD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                dst[i * zstride<P, NS>()] = x[q * RAD + i];
            });
        }
    });
    group_sync<P, NTHREADS>(grp);
}

// Full forward FFT of one symbol.  On entry x[i] = z[t + i*T] (i < R); on exit the natural-order
// spectrum Z[0..M) sits in zs (padded indexing, or plain indexing when NATURAL) and is visible to
// the symbol's T threads.
template <class P, int NTHREADS, bool NATURAL = false>
__device__ __forceinline__ void fft_forward(float2 (&x)[P::R], float2* __restrict__ zs,
                                            const float2* __restrict__ tw, int t, int grp) {
    fft_pass<P, NTHREADS, 0>(x, zs, tw, t, grp);
    if constexpr (P::NPASS > 2) {
        fft_pass<P, NTHREADS, 1>(x, zs, tw, t, grp);
        fft_pass<P, NTHREADS, 2, NATURAL>(x, zs, tw, t, grp);
    } else {
        fft_pass<P, NTHREADS, 1, NATURAL>(x, zs, tw, t, grp);
    }
}

// The same transform with the last pass left in registers (see fft_pass<.., STORE = false>).
template <class P, int NTHREADS>
__device__ __forceinline__ void fft_forward_to_regs(float2 (&x)[P::R], float2* __restrict__ zs,
                                                    const float2* __restrict__ tw, int t, int grp) {
    fft_pass<P, NTHREADS, 0>(x, zs, tw, t, grp);
    if constexpr (P::NPASS > 2) {
        fft_pass<P, NTHREADS, 1>(x, zs, tw, t, grp);
        fft_pass<P, NTHREADS, 2, true, false>(x, zs, tw, t, grp);
    } else {
        fft_pass<P, NTHREADS, 1, true, false>(x, zs, tw, t, grp);
    }
}

// Host-side: fill the twiddle table of plan P (double precision, rounded to float).
template <class P>
inline void fill_twiddles(float2* out) {
    for (int pass = 1; pass < P::NPASS; ++pass) {
        const int RAD = P::rad(pass), NS = P::ns(pass), Q = P::R / RAD;
        float2* o = out + P::tw_off(pass);
        if (pass == 1 && P::DEDUP1) {
            for (int i = 1; i < RAD; ++i)
                for (int c = 0; c < NS; ++c) {
                    const double ang = -2.0 * 3.14159265358979323846 * (double)(c * i) / (double)(NS * RAD);
                    o[(i - 1) * NS + c] = make_float2((float)cos(ang), (float)sin(ang));
                }
            continue;
        }
        const bool pair = P::PAIRED || (P::PAIRLAST && pass == P::NPASS - 1);
        for (int q = 0; q < Q; ++q)
            for (int i = 1; i < RAD; ++i)
                for (int t = 0; t < P::T; ++t) {
                    const int j = pair ? 2 * t + q : t + q * P::T;
                    const double ang = -2.0 * 3.14159265358979323846 * (double)((j % NS) * i) / (double)(NS * RAD);
                    if (P::TWAB) {            // float4 (wr, wi, -wi, wr) at [(i-1)][q][t]
                        const int idx4 = ((i - 1) * 2 + q) * P::T + t;
                        o[2 * idx4] = make_float2((float)cos(ang), (float)sin(ang));
                        o[2 * idx4 + 1] = make_float2(-(float)sin(ang), (float)cos(ang));
                        continue;
                    }
                    const int idx = pair ? (i - 1) * 2 * P::T + 2 * t + q : (q * (RAD - 1) + (i - 1)) * P::T + t;
                    o[idx] = make_float2((float)cos(ang), (float)sin(ang));
                }
    }
}

}  // namespace gf3
