#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float2 a[8], b = make_float2(s, s * 1.0001f), c = make_float2(0.5f * s, 0.25f * s);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }   // 2 scalar FFMA
            else if (MODE == 1) a[i] = ffma2(a[i], b, c);                                          // 1 FFMA2
            else if (MODE == 2) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }                  // 2 FADD
            else a[i] = fadd2(a[i], b);                                                            // 1 FADD2
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> float run(float* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 10, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    int iters = 20000;
    double flop_pairs = 148.0 * 8 * 256 * iters * 8;   // float2 ops
    const char* names[4] = {"2xFFMA", "FFMA2", "2xFADD", "FADD2"};
    float ms[4] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters)};
    for (int m = 0; m < 4; ++m) printf("%-7s %8.3f ms  %.1f G float2-ops/s  (%.1f T lane-ops/s)\n", names[m], ms[m], flop_pairs / ms[m] / 1e6, 2 * flop_pairs / ms[m] / 1e9);
    return 0;
}
