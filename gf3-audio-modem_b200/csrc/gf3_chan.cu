// gf3_chan.cu -- synthetic-channel and bit-error kernels for the batched sweeps.
//
//   channel_sim_kernel : y = lfilter(taps, 1, x) + sigma * N(0,1)   (per-stream FIR, Philox noise)
//   ber_count_kernel   : popcount(a ^ b) -> u64 counters (all-reduced over NCCL by the host)
#include "gf3_common.cuh"

namespace gf3 {

// Philox-4x32-10 (Salmon et al. 2011), counter-based: reproducible for any grid shape.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;    // (0, 1]
    const float u2 = (float)b * 2.3283064365386963e-10f;             // [0, 1)
    const float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(rad * c, rad * s);
}

constexpr int kMaxTaps = 64;

struct ChanArgs {
    const float* x;
    const float* taps;
    const float* sigma;
    const int64_t* ids;         // optional: Philox stream id of every row (default: the row index)
    float* y;
    int64_t x_stride, y_stride, T;
    uint64_t seed;
    int n_taps;
};

// grid.y = stream; a CTA walks tiles of 1024 consecutive outputs; each thread produces 4 of them.
// The tile's input window (tile + n_taps - 1 samples of history, zero outside [0, T)) is staged in
// shared memory; a thread slides an 8-sample register window over it four taps at a time, so every
// multiply-add takes its operands from registers (128-bit conflict-free shared-memory reads only).
// Summation order (k ascending, one fused multiply-add per tap) and the Philox counter per group of
// four samples are those of the straightforward loop, so results do not depend on the tiling.
__global__ void __launch_bounds__(256) channel_sim_kernel(const ChanArgs a) {
    constexpr int NT = 256, TILE = NT * 4, HIST = kMaxTaps;              // HIST: multiple of 4, >= n_taps - 1
    __shared__ __align__(16) float h[kMaxTaps];
    __shared__ __align__(16) float xs[HIST + TILE];
    const int tid = threadIdx.x;
    const int64_t stream = blockIdx.y;
    if (tid < kMaxTaps) h[tid] = tid < a.n_taps ? a.taps[stream * a.n_taps + tid] : 0.f;
    const float* x = a.x + stream * a.x_stride;
    float* y = a.y + stream * a.y_stride;
    const float sg = a.sigma ? a.sigma[stream] : 0.f;
    const int64_t sid = a.ids ? a.ids[stream] : stream;                  // the noise of a row depends on its id only
    const int nblk = (a.n_taps + 3) >> 2;                                // tap blocks of four
    const bool y_al16 = ((reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    for (int64_t tile0 = (int64_t)blockIdx.x * TILE; tile0 < a.T; tile0 += (int64_t)gridDim.x * TILE) {
        __syncthreads();                                                  // previous tile fully consumed (and h written)
        for (int i = tid; i < HIST + TILE; i += NT) {
            const int64_t n = tile0 - HIST + i;
            xs[i] = (n >= 0 && n < a.T) ? x[n] : 0.f;
        }
        __syncthreads();
        const int64_t n0 = tile0 + 4 * tid;
        if (n0 < a.T) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            // taps k = 4 kb + kk use samples n0 + e - k = (n0 - 4 kb) + (e - kk): the window w[0..7] = xs at
            // (4 tid + HIST - 4 kb) - 4 .. + 3 covers e - kk in [-3, 3]
            const float4* xw = reinterpret_cast<const float4*>(xs) + tid + HIST / 4;
            const float4* hw = reinterpret_cast<const float4*>(h);
            for (int kb = 0; kb < nblk; ++kb) {
                const float4 lo = xw[-kb - 1], hi = xw[-kb], hk = hw[kb];
                const float w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                const float hv[4] = {hk.x, hk.y, hk.z, hk.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] = fmaf(hv[kk], w[4 + e - kk], acc[e]);
            }
            if (sg != 0.f) {
                const int64_t q = n0 >> 2;
                const uint4 rnd = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)sid, (uint32_t)(sid >> 32)),
                                                make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                const float2 g0 = box_muller(rnd.x, rnd.y), g1 = box_muller(rnd.z, rnd.w);
                acc[0] = fmaf(sg, g0.x, acc[0]); acc[1] = fmaf(sg, g0.y, acc[1]);
                acc[2] = fmaf(sg, g1.x, acc[2]); acc[3] = fmaf(sg, g1.y, acc[3]);
            }
            if (y_al16 && (a.y_stride & 3) == 0 && n0 + 3 < a.T) {
                *reinterpret_cast<float4*>(y + n0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (n0 + e < a.T) y[n0 + e] = acc[e];
            }
        }
    }
}

__global__ void __launch_bounds__(256) ber_count_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                       int64_t nbits, unsigned long long* counter) {
    const int64_t nbytes = nbits >> 3;
    unsigned long long errs = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (int64_t)gridDim.x * blockDim.x)
        errs += __popc((unsigned)(a[i] ^ b[i]));
    if (blockIdx.x == 0 && threadIdx.x == 0 && (nbits & 7)) {
        const unsigned mask = (0xFF00u >> (nbits & 7)) & 0xFFu;       // leading (MSB-first) bits of the last byte
        errs += __popc((unsigned)((a[nbytes] ^ b[nbytes]) & mask));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) errs += __shfl_xor_sync(0xffffffffu, errs, o);
    __shared__ unsigned long long wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = errs;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        if (t) atomicAdd(counter, t);
        if (blockIdx.x == 0) atomicAdd(counter + 1, (unsigned long long)nbits);
    }
}

// Uniform random bytes, 16 per Philox call; row r draws from the counter stream of row_ids[r] (default r),
// so a row's bytes depend on (seed, id, position) only -- not on how rows are batched or sharded.
__global__ void __launch_bounds__(256) random_bytes_kernel(uint8_t* __restrict__ out, int64_t out_stride, int64_t row_bytes,
                                                          const int64_t* __restrict__ row_ids, uint64_t seed) {
    const int64_t row = blockIdx.y;
    const int64_t id = row_ids ? row_ids[row] : row;
    uint8_t* o = out + row * out_stride;
    const int64_t n16 = (row_bytes + 15) >> 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n16; q += (int64_t)gridDim.x * blockDim.x) {
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)id, (uint32_t)(id >> 32) ^ 0x52424e44u),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        if (16 * q + 16 <= row_bytes && ((reinterpret_cast<uintptr_t>(o) | (uintptr_t)out_stride) & 15) == 0) {
            reinterpret_cast<uint4*>(o)[q] = rnd;
        } else {
            for (int e = 0; e < 16 && 16 * q + e < row_bytes; ++e) o[16 * q + e] = (uint8_t)(w[e >> 2] >> (8 * (e & 3)));
        }
    }
}

// PCM samples -> float32 (exact value conversion), 16 samples per thread step, 128-bit stores
template <typename T>
__global__ void __launch_bounds__(256) pcm_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n && (reinterpret_cast<uintptr_t>(out + i) & 15) == 0) {
            float4 v = make_float4((float)in[i], (float)in[i + 1], (float)in[i + 2], (float)in[i + 3]);
            *reinterpret_cast<float4*>(out + i) = v;
        } else {
            for (int e = 0; e < 4 && i + e < n; ++e) out[i + e] = (float)in[i + e];
        }
    }
}

}  // namespace gf3

using namespace gf3;

extern "C" int gf3_pcm_to_f32(const void* pcm, int32_t format, int64_t n, float* out, void* stream) {
    GF3_REQUIRE(pcm && out, "pcm_to_f32: null argument");
    GF3_REQUIRE(format == 0 || format == 1, "pcm_to_f32: format must be 0 (uint8) or 1 (int16)");
    GF3_REQUIRE(n >= 0, "pcm_to_f32: negative length");
    if (n == 0) return GF3_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (format == 0) pcm_to_f32_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(pcm), out, n);
    else pcm_to_f32_kernel<int16_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const int16_t*>(pcm), out, n);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

static int channel_sim_common(const float* x, int64_t x_stride, int64_t n_streams, int64_t T,
                              const float* taps, int32_t n_taps, const float* sigma, const int64_t* ids, uint64_t seed,
                              float* y, int64_t y_stride, void* stream) {
    GF3_REQUIRE(x && taps && y, "channel_sim: null argument");
    GF3_REQUIRE(n_taps >= 1 && n_taps <= kMaxTaps, "channel_sim: n_taps must be in 1..%d", kMaxTaps);
    GF3_REQUIRE(n_streams >= 0 && n_streams <= 65535 && T >= 0, "channel_sim: bad sizes (n_streams <= 65535)");
    GF3_REQUIRE(x != y, "channel_sim: in-place operation is not supported");
    if (n_streams == 0 || T == 0) return GF3_OK;
    ChanArgs a;
    a.x = x; a.taps = taps; a.sigma = sigma; a.ids = ids; a.y = y; a.x_stride = x_stride; a.y_stride = y_stride; a.T = T;
    a.seed = seed; a.n_taps = n_taps;
    int64_t gx = (T + 1023) / 1024;                  // tiles of 1024 outputs per stream
    const int64_t cap = (148LL * 8 * 4 + n_streams - 1) / n_streams;      // a few waves in total
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    channel_sim_kernel<<<dim3((unsigned)gx, (unsigned)n_streams), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_channel_sim(const float* x, int64_t x_stride, int64_t n_streams, int64_t T,
                               const float* taps, int32_t n_taps, const float* sigma, uint64_t seed,
                               float* y, int64_t y_stride, void* stream) {
    return channel_sim_common(x, x_stride, n_streams, T, taps, n_taps, sigma, nullptr, seed, y, y_stride, stream);
}

extern "C" int gf3_channel_sim_ids(const float* x, int64_t x_stride, int64_t n_streams, int64_t T,
                                   const float* taps, int32_t n_taps, const float* sigma, const int64_t* stream_ids,
                                   uint64_t seed, float* y, int64_t y_stride, void* stream) {
    GF3_REQUIRE(stream_ids != nullptr, "channel_sim_ids: null stream_ids");
    return channel_sim_common(x, x_stride, n_streams, T, taps, n_taps, sigma, stream_ids, seed, y, y_stride, stream);
}

extern "C" int gf3_random_bytes(uint8_t* out, int64_t out_stride, int64_t n_rows, int64_t row_bytes,
                                const int64_t* row_ids, uint64_t seed, void* stream) {
    GF3_REQUIRE(out != nullptr, "random_bytes: null output");
    GF3_REQUIRE(n_rows >= 0 && n_rows <= 65535 && row_bytes >= 0 && out_stride >= row_bytes, "random_bytes: bad sizes (n_rows <= 65535)");
    if (n_rows == 0 || row_bytes == 0) return GF3_OK;
    int64_t gx = (row_bytes + 16 * 256 - 1) / (16 * 256);
    if (gx > 64) gx = 64;
    random_bytes_kernel<<<dim3((unsigned)gx, (unsigned)n_rows), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        out, out_stride, row_bytes, row_ids, seed);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_ber_count(const uint8_t* a, const uint8_t* b, int64_t nbits, uint64_t* counter, void* stream) {
    GF3_REQUIRE(a && b && counter, "ber_count: null argument");
    GF3_REQUIRE(nbits >= 0, "ber_count: negative bit count");
    if (nbits == 0) return GF3_OK;
    int64_t blocks = ((nbits >> 3) + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    ber_count_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, nbits, reinterpret_cast<unsigned long long*>(counter));
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}
