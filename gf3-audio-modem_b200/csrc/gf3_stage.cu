// gf3_stage.cu -- stage-level entry points of the receive / transmit chain (sm_100a).
//
// The fused kernels in gf3_rx.cu / gf3_tx.cu cover receiver.receive() and transmitter.transmit().
// The reference also exposes the individual stages as public methods that take and return
// spectra (receiver.equalise, receiver.demap, transmitter.send_to_stream; OFDM.py:422-480,
// 484-500, 242-276).  These kernels run those stages on arrays the caller already holds in
// the frequency / time domain, with the same arithmetic as the fused path.
#include "gf3_common.cuh"
#include "gf3_fft.cuh"
#include "gf3_fit.cuh"

namespace gf3 {

__device__ __forceinline__ float2 st_expj(double a) {      // exp(+j a): reduced in double, evaluated in float
    const double inv2pi = 0.15915494309189533577;
    double r = a * inv2pi;
    r -= rint(r);
    float s, c;
    sincospif(2.0f * (float)r, &s, &c);
    return make_float2(c, s);
}

// ---------------------------------------------------------------- equalise, first half (OFDM.py:429-462)
// One CTA per packet: Hs = mean_P(start)/known, He = mean_P(end)/known, phases of the fit window,
// unwrap + least-squares slope (gf3_fit.cuh).
struct EqEstArgs {
    const float2* start;     // [n_packets, P, K]
    const float2* end;       // [n_packets, P, K]
    const float2* known;     // [K]
    float2* Hs;              // [n_packets, K]
    float2* He;
    double* slope;           // [n_packets]
    int K, P, fit_lo, fit_hi;
};

__global__ void __launch_bounds__(256) eq_estimate_kernel(const EqEstArgs a) {
    constexpr int NT = 256;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* phi = reinterpret_cast<double*>(smem_raw);             // [2][K]
    __shared__ int warp_tot[NT / 32];
    __shared__ double red[NT / 32];
    const int tid = threadIdx.x, K = a.K;
    const int64_t pkt = blockIdx.x;
    const int flo = max(0, min(a.fit_lo, K)), fhi = max(flo, min(a.fit_hi, K));
    const float invP = 1.0f / (float)a.P;
    for (int item = tid; item < 2 * K; item += NT) {
        const int blk = item / K, n = item - blk * K;
        const float2* src = (blk ? a.end : a.start) + pkt * (int64_t)a.P * K + n;
        float2 acc = make_float2(0.f, 0.f);
        for (int p = 0; p < a.P; ++p) {
            const float2 v = src[(int64_t)p * K];
            acc.x += v.x;
            acc.y += v.y;
        }
        const float2 kn = a.known[n];                                // |known| = 1: 1/known = conj(known)
        float2 h = cmul(make_float2(acc.x * invP, acc.y * invP), cconj(kn));
        ((blk ? a.He : a.Hs) + pkt * K)[n] = h;
        if (n >= flo && n < fhi) phi[blk * K + n] = atan2((double)h.y, (double)h.x);
    }
    __syncthreads();
    const double sl = fit_slope<NT>(phi, K, flo, fhi, warp_tot, red);
    if (tid == 0) a.slope[pkt] = sl;
}

// ---------------------------------------------------------------- equalise, second half (OFDM.py:466-478)
// Elementwise over [n_packets, L, K]: Hest = (|Hs| + (|He|-|Hs|) w) exp(j (angle Hs + p n w)),
// w = (l + P/2)/(L + P), n the 0-based bin index; data_eq = data / Hest.
struct EqApplyArgs {
    const float2* data;      // [n_packets, L, K]
    const float2* Hs;
    const float2* He;
    const double* slope;
    float2* eq;              // [n_packets, L, K]
    float2* Hest;            // [n_packets, L, K] or null
    int K, P, L;
    int64_t total;           // n_packets * L * K
};

__global__ void __launch_bounds__(256) eq_apply_kernel(const EqApplyArgs a) {
    const double inv_lp = 1.0 / (double)(a.L + a.P);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % a.K);
        const int64_t pl = i / a.K;
        const int l = (int)(pl % a.L);
        const int64_t pkt = pl / a.L;
        const float2 hs = a.Hs[pkt * a.K + n], he = a.He[pkt * a.K + n];
        const double w = ((double)l + 0.5 * (double)a.P) * inv_lp;
        const float as = sqrtf(hs.x * hs.x + hs.y * hs.y), ae = sqrtf(he.x * he.x + he.y * he.y);
        const float mag = as + (ae - as) * (float)w;
        // exp(j angle(Hs)) = Hs/|Hs|  (angle(0) = 0 in numpy)
        const float2 u = as > 0.f ? make_float2(hs.x / as, hs.y / as) : make_float2(1.f, 0.f);
        const float2 rot = cmul(u, st_expj(a.slope[pkt] * (double)n * w));
        const float2 H = make_float2(mag * rot.x, mag * rot.y);
        if (a.Hest) a.Hest[i] = H;
        // data / H = data * conj(rot) / mag
        const float2 d = a.data[i];
        const float2 q = cmul(d, cconj(rot));
        a.eq[i] = make_float2(q.x / mag, q.y / mag);
    }
}

// ---------------------------------------------------------------- demap (OFDM.py:484-500)
// Minimum distance over [(0,0),(1,0),(1,1),(0,1)] in that order == b0 = imag < 0, b1 = real < 0; on an
// exact tie (a coordinate == 0) argmin keeps the FIRST minimum of that order: b1 = 0, and b0 = 1 only for
// imag == 0 with real < 0 (pinned against the literal arg-min form by the golden tests).
__global__ void __launch_bounds__(256) demap_kernel(const float2* sym, uint8_t* bits, float2* hard, int64_t n) {
    const float h = 0.70710678118654752440f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 y = sym[i];
        const unsigned b1 = y.x < 0.f, b0 = (y.y < 0.f) || (y.y == 0.f && b1);
        reinterpret_cast<uchar2*>(bits)[i] = make_uchar2((unsigned char)b0, (unsigned char)b1);
        if (hard) hard[i] = make_float2(b1 ? -h : h, b0 ? -h : h);
    }
}

// ---------------------------------------------------------------- send_to_stream: data symbols (OFDM.py:251-257)
struct FrameDataArgs {
    const float* data;       // [n_packets, L*symlen]
    float* out;
    int64_t pkt_len, data_off, per_pkt, total;   // per_pkt = L*symlen, data_off = sync_len + P*symlen
    float gain;
};
__global__ void __launch_bounds__(256) frame_data_kernel(const FrameDataArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pk = i / a.per_pkt, r = i - pk * a.per_pkt;
        a.out[pk * a.pkt_len + a.data_off + r] = a.gain * a.data[i];
    }
}

static unsigned grid_for(int64_t total, int sm_count) {
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// Y / H, H broadcast over the rows (the old API's module-level equalise(Y, H), Weekend Challenge.ipynb:225)
__global__ void __launch_bounds__(256) cdiv_kernel(const float2* __restrict__ Y, const float2* __restrict__ H, int64_t n_rows, int m,
                                                   float2* __restrict__ out) {
    const int64_t total = n_rows * m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 y = Y[i], h = H[i % m];
        const float d = 1.0f / (h.x * h.x + h.y * h.y);
        out[i] = make_float2((y.x * h.x + y.y * h.y) * d, (y.y * h.x - y.x * h.y) * d);
    }
}

int tx_frame_known(const gf3_plan* plan, const float* known, const float* sync, int sync_len, int64_t n_packets,
                   float* out, cudaStream_t st);     // gf3_tx.cu

}  // namespace gf3

using namespace gf3;

extern "C" int gf3_eq_estimate(const gf3_plan* plan, const float* start, const float* end, int64_t n_packets,
                               const float* known, float* Hs, float* He, double* slope, void* stream) {
    GF3_REQUIRE(plan && start && end && known && Hs && He && slope, "eq_estimate: null argument");
    GF3_REQUIRE(n_packets >= 0 && n_packets <= 0x7fffffff, "eq_estimate: bad packet count");
    const gf3_params& p = plan->p;
    GF3_REQUIRE(p.n_pilots >= 1, "eq_estimate: n_pilots must be >= 1 (OFDM.py:424 short-circuits no_pilots == 0)");
    if (n_packets == 0) return GF3_OK;
    EqEstArgs a;
    a.start = reinterpret_cast<const float2*>(start); a.end = reinterpret_cast<const float2*>(end);
    a.known = reinterpret_cast<const float2*>(known);
    a.Hs = reinterpret_cast<float2*>(Hs); a.He = reinterpret_cast<float2*>(He); a.slope = slope;
    a.K = p.N / 2 - 1; a.P = p.n_pilots; a.fit_lo = p.fit_lo; a.fit_hi = p.fit_hi;
    const size_t smem = (size_t)2 * a.K * sizeof(double);
    GF3_CHECK_CUDA(cudaFuncSetAttribute(eq_estimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eq_estimate_kernel<<<(unsigned)n_packets, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_eq_apply(const gf3_plan* plan, const float* data, int64_t n_packets, const float* Hs,
                            const float* He, const double* slope, float* eq, float* Hest, void* stream) {
    GF3_REQUIRE(plan && data && Hs && He && slope && eq, "eq_apply: null argument");
    GF3_REQUIRE(n_packets >= 0, "eq_apply: negative packet count");
    if (n_packets == 0) return GF3_OK;
    const gf3_params& p = plan->p;
    EqApplyArgs a;
    a.data = reinterpret_cast<const float2*>(data); a.Hs = reinterpret_cast<const float2*>(Hs);
    a.He = reinterpret_cast<const float2*>(He); a.slope = slope;
    a.eq = reinterpret_cast<float2*>(eq); a.Hest = reinterpret_cast<float2*>(Hest);
    a.K = p.N / 2 - 1; a.P = p.n_pilots; a.L = p.packet_len;
    a.total = n_packets * (int64_t)a.L * a.K;
    if (a.total == 0) return GF3_OK;
    eq_apply_kernel<<<grid_for(a.total, plan->sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_demap(const float* symbols, int64_t n, uint8_t* bits, float* hard, void* stream) {
    GF3_REQUIRE(n >= 0, "demap: negative count");
    if (n == 0) return GF3_OK;
    GF3_REQUIRE(symbols && bits, "demap: null argument");
    GF3_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 1) == 0, "demap: bits must be 2-byte aligned");
    int sms = 148, dev = 0;
    GF3_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    demap_kernel<<<grid_for(n, sms), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(symbols), bits, reinterpret_cast<float2*>(hard), n);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_tx_frame(const gf3_plan* plan, const float* data_time, int64_t n_packets, const float* sync,
                            int32_t sync_len, const float* known, float* out, void* stream) {
    GF3_REQUIRE(plan && sync && known && out, "tx_frame: null argument");
    GF3_REQUIRE(n_packets >= 0 && sync_len >= 0, "tx_frame: negative size");
    GF3_REQUIRE(n_packets == 0 || data_time, "tx_frame: null data");
    const gf3_params& p = plan->p;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int symlen = p.N + p.cp;
    int rc = tx_frame_known(plan, known, sync, sync_len, n_packets, out, st);
    if (rc) return rc;
    FrameDataArgs a;
    a.data = data_time; a.out = out;
    a.pkt_len = (int64_t)sync_len + (int64_t)(2 * p.n_pilots + p.packet_len) * symlen;
    a.data_off = (int64_t)sync_len + (int64_t)p.n_pilots * symlen;
    a.per_pkt = (int64_t)p.packet_len * symlen;
    a.total = n_packets * a.per_pkt;
    a.gain = p.tx_gain;
    if (a.total > 0) {
        frame_data_kernel<<<grid_for(a.total, plan->sm_count), 256, 0, st>>>(a);
        GF3_LAUNCH_CHECK();
    }
    return GF3_OK;
}

extern "C" int gf3_cdiv(const float* Y, const float* H, int64_t n_rows, int32_t m, float* out, void* stream) {
    GF3_REQUIRE(Y && H && out, "cdiv: null argument");
    GF3_REQUIRE(n_rows >= 0 && m >= 1, "cdiv: bad sizes");
    if (n_rows == 0) return GF3_OK;
    int64_t blocks = (n_rows * m + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cdiv_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(Y), reinterpret_cast<const float2*>(H),
                                                                                   n_rows, m, reinterpret_cast<float2*>(out));
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}
