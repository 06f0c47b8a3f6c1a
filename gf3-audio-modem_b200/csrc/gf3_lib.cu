// gf3_lib.cu -- library plumbing: error strings, parameter defaults, the plan handle.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "gf3_common.cuh"
#include "gf3_fft.cuh"

namespace gf3 {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <class P>
static void fill_tw_vec(std::vector<float2>& v) {
    v.assign(P::TW_TOTAL > 0 ? P::TW_TOTAL : 1, make_float2(0.f, 0.f));
    fill_twiddles<P>(v.data());
}

static bool host_twiddles(int logN, std::vector<float2>& v) {
    switch (logN) {
        case 6: fill_tw_vec<FftPlan<6>>(v); return true;
        case 7: fill_tw_vec<FftPlan<7>>(v); return true;
        case 8: fill_tw_vec<FftPlan<8>>(v); return true;
        case 9: fill_tw_vec<FftPlan<9>>(v); return true;
        case 10: fill_tw_vec<FftPlan<10>>(v); return true;
        case 11: fill_tw_vec<FftPlan<11>>(v); return true;
        case 12: fill_tw_vec<FftPlan<12>>(v); return true;
        case 112: fill_tw_vec<FftPlan12B>(v); return true;        // N = 4096 as 32 x 8 x 8 (data-symbol kernel)
        case 212: fill_tw_vec<FftPlan12C>(v); return true;        // N = 4096 as 64 x 32
        case 312: fill_tw_vec<FftPlan12P>(v); return true;        // N = 4096 as 16 x 16 x 8, last pass paired
        default: return false;
    }
}

int upload_twiddles(int logN, float2** d_out) {
    std::vector<float2> h;
    GF3_REQUIRE(host_twiddles(logN, h), "unsupported FFT size 2^%d (supported: 64..4096)", logN);
    GF3_CHECK_CUDA(cudaMalloc(d_out, h.size() * sizeof(float2)));
    GF3_CHECK_CUDA(cudaMemcpy(*d_out, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return GF3_OK;
}

int sync_plan_init(gf3_plan* plan);     // gf3_sync.cu
void sync_plan_free(gf3_plan* plan);

}  // namespace gf3

using namespace gf3;

extern "C" int gf3_abi_version(void) { return GF3_ABI_VERSION; }

extern "C" const char* gf3_last_error(void) { return g_err; }

extern "C" int64_t gf3_launch_count(void) { return g_launches.load(); }

extern "C" int gf3_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gf3_params_default(gf3_params* p, int N, int cp, int lo, int hi, int n_pilots, int packet_len) {
    GF3_REQUIRE(p != nullptr, "params_default: null");
    memset(p, 0, sizeof(*p));
    p->N = N; p->cp = cp; p->lo = lo; p->hi = hi; p->n_pilots = n_pilots; p->packet_len = packet_len;
    p->fit_lo = 500; p->fit_hi = 1000;            // OFDM.py:462
    p->chirp_len = 5 * (N + cp);                  // OFDM.py:64
    p->fs = 48000.f; p->f0 = 0.f; p->f1 = 8000.f; // OFDM.py:24,62,63
    p->thresh = 0.4f;                             // OFDM.py:361
    p->tx_gain = 2.0f;                            // OFDM.py:256
    p->chirp_gain = 0.2f;                         // OFDM.py:109
    return GF3_OK;
}

static int validate(const gf3_params& p) {
    GF3_REQUIRE(p.N >= 64 && p.N <= 4096 && (p.N & (p.N - 1)) == 0, "N = %d must be a power of two in 64..4096", p.N);
    GF3_REQUIRE(p.cp >= 0, "cp = %d must be >= 0", p.cp);
    GF3_REQUIRE(p.lo >= 1 && p.hi > p.lo && p.hi <= p.N / 2, "data bins [%d, %d) must satisfy 1 <= lo < hi <= N/2", p.lo, p.hi);
    GF3_REQUIRE(p.n_pilots >= 0 && p.packet_len >= 1, "n_pilots >= 0 and packet_len >= 1 required");
    GF3_REQUIRE(p.chirp_len >= 2, "chirp_len must be >= 2");
    GF3_REQUIRE(p.fs > 0.f, "fs must be positive");
    return GF3_OK;
}

extern "C" int gf3_plan_create(const gf3_params* p, gf3_plan** out) {
    GF3_REQUIRE(p && out, "plan_create: null argument");
    *out = nullptr;
    int rc = validate(*p);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libgf3b200 has no CPU path (%s)", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return GF3_ERR_NODEVICE;
    }
    gf3_plan* plan = new gf3_plan();
    memset(plan, 0, sizeof(*plan));
    plan->p = *p;
    plan->logN = ilog2(p->N);
    // every failure below goes through gf3_plan_destroy: the handle is zero-initialised, so whatever was
    // allocated so far is released and nothing else is touched
    auto fill = [&]() -> int {
        GF3_CHECK_CUDA(cudaGetDevice(&plan->device));
        GF3_CHECK_CUDA(cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, plan->device));
        int r = upload_twiddles(plan->logN, &plan->d_tw);
        if (r) return r;
        plan->d_tw_demod = plan->d_tw;
#if GF3_RX12_ALT
        if (plan->logN == 12) { r = upload_twiddles(GF3_RX12_ALT * 100 + 12, &plan->d_tw_demod); if (r) return r; }
#endif
        const int K = p->N / 2 - 1;
        std::vector<float2> ones(K, make_float2(1.f, 0.f));
        GF3_CHECK_CUDA(cudaMalloc(&plan->d_ones, K * sizeof(float2)));
        GF3_CHECK_CUDA(cudaMemcpy(plan->d_ones, ones.data(), K * sizeof(float2), cudaMemcpyHostToDevice));
        return sync_plan_init(plan);
    };
    rc = fill();
    if (rc) { gf3_plan_destroy(plan); return rc; }
    *out = plan;
    return GF3_OK;
}

extern "C" int gf3_plan_destroy(gf3_plan* plan) {
    if (!plan) return GF3_OK;
    sync_plan_free(plan);
    if (plan->d_tw_demod && plan->d_tw_demod != plan->d_tw) cudaFree(plan->d_tw_demod);
    if (plan->d_tw) cudaFree(plan->d_tw);
    if (plan->d_ones) cudaFree(plan->d_ones);
    delete plan;
    return GF3_OK;
}

extern "C" int gf3_plan_params(const gf3_plan* plan, gf3_params* out) {
    GF3_REQUIRE(plan && out, "plan_params: null argument");
    *out = plan->p;
    return GF3_OK;
}
