// gf3_common.cuh -- shared host-side plumbing of libgf3b200.so: error reporting, the plan
// handle, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/gf3_b200.h"

namespace gf3 {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define GF3_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            gf3::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,  \
                           __LINE__);                                                          \
            return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)            \
                       ? GF3_ERR_NODEVICE                                                      \
                       : GF3_ERR_CUDA;                                                         \
        }                                                                                      \
    } while (0)

#define GF3_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            gf3::set_error(__VA_ARGS__);                                                       \
            return GF3_ERR_INVALID;                                                            \
        }                                                                                      \
    } while (0)

#define GF3_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        gf3::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        GF3_CHECK_CUDA(cudaGetLastError());                                                    \
    } while (0)

static inline int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace gf3

// Plan of the N = 4096 data-symbol kernel: 3 = 16 x 16 x 8 on 128 threads per symbol with the last pass on adjacent
// columns (128-bit exchange; C4 0.604 vs 0.607 ms), 0 = the same with 64-bit accesses throughout (the plan of the
// matched filter and the estimate kernel), 1 = 32 x 8 x 8 on 64 threads (0.759), 2 = 64 x 32 on one warp (0.749)
#ifndef GF3_RX12_ALT
#define GF3_RX12_ALT 3
#endif
#ifndef GF3_FUSE12
#define GF3_FUSE12 0           // with the 32 x 8 x 8 plan both pilot blocks fit the spectrum buffer: fuse the estimate at N = 4096 too?
#endif

// The opaque handle of include/gf3_b200.h.
struct gf3_plan {
    gf3_params p;
    int logN;
    int device;
    int sm_count;
    float2* d_tw;        // FFT twiddle table of the N-point symbol plan
    float2* d_tw_demod;  // twiddle table of the data-symbol kernel's plan (== d_tw unless that kernel uses another factorisation)
    float2* d_ones;      // K ones (unit channel for gf3_rx_spectrum)
    // sync (overlap-save matched filter): block FFT size NB real samples, hop HB
    int sync_logN;       // log2 of the overlap-save FFT length
    int sync_parts;      // filter partitions
    float2* d_sync_tw;   // twiddles of the sync FFT plan
    float2* d_chirp_spec;  // [sync_parts][NB/2+1] spectrum of the time-reversed chirp partitions
    float* d_chirp;      // [chirp_len] sync chirp
    float* d_chirp_pairs;  // [sync_parts][8][128] float4: the partitions in the fused matched filter's bin-pair layout, pre-scaled
    float2* d_chirp_dc;    // [sync_parts] (H[0], H[M]) pre-scaled
    float2* d_chirp_one;   // [NB/2] the unit partition (inverse stage of the three-kernel matched filter for long chirps)
    float* d_chirp_energy; // [sync_parts][32] weighted energy of every partition's spectrum per bin group (detection-only bound)
};
